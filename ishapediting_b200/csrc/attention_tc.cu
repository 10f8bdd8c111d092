// attention_tc.cu — fused self-attention FORWARD on the 5th-generation tensor cores (tcgen05.mma, operands fed by TMA,
// S and O accumulators in TMEM) for the layers with T % 128 == 0 tokens (32x32 and 16x16 resolution: T = 1024, 256).
// Reference: guided_diffusion/unet.py:337-354 (QKVAttentionLegacy.forward, 64 channels per head, softmax in fp32).
//
// One CTA per (128 queries, head, image), 192 threads:
//   warp 5   TMA producer: Q once, K / V tiles of 128 keys through two-slot rings (box {64 ch, 128 tokens} of the
//            [N*T, 3C] qkv matrix at the head's q / k / v channel block — the legacy interleave of unet.py:346)
//   warp 4   MMA issuer:   S = Q K^T (M 128, N 128, K 64) into one of two TMEM buffers;  O += P V (M 128, N 64, K 128)
//            with V as an MN-major B operand straight from the TMA image (no transpose pass)
//   warps 0-3 softmax:     one query row per thread, read from TMEM with tcgen05.ld
// Two passes over the key blocks instead of an online softmax: pass 1 only takes the row maxima (S = Q K^T is cheap on
// UMMA and T <= 1024), pass 2 recomputes S, writes P = exp2(S - max) as the bf16 A operand of the PV product and sums
// the row — O never needs rescaling, so nothing in TMEM is read-modify-written and QK^T of block j+1 runs under the
// exponentials of block j.  Outputs: O (bf16) and the row log-sum-exp (log2 domain), exactly what fa_fwd_kernel
// writes, so the mma.sync backward (attention_flash.cu) consumes either.
#include "common.cuh"

namespace isb {
namespace fatc {

constexpr int BM = 128, BN = 128, D = 64;
constexpr int THREADS = 192;
constexpr int TILE = 128 * 128;              // 128 rows x 64 bf16 = 16 KiB (Q, K, V tiles; one K-atom of P)
constexpr int OFF_Q = 0, OFF_K = TILE, OFF_V = 3 * TILE, OFF_P = 5 * TILE;
constexpr int SMEM_BYTES = 7 * TILE + 1024 + 4096;  // Q + 2 K + 2 V + P (2 atoms) + alignment slack; > half an SM: one CTA
                                                    // per SM, so that two CTAs never queue for the 512 TMEM columns

struct Params {
  __nv_bfloat16* out;
  float* lse;
  int T, heads;
  float scale_log2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {     // bounded: a bug traps, never hangs
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 6000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (K-major A / B, and MN-major B: see header)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]),
      "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]),
      "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of the 16-byte chunk holding keys [k8*8, k8*8+8) of query row r inside the 128 x 128 bf16 P operand
__device__ __forceinline__ uint32_t p_chunk_off(int r, int k8) {
  return static_cast<uint32_t>((k8 >> 3) * TILE + (r >> 3) * 1024 + (r & 7) * 128 + (((k8 & 7) ^ (r & 7)) << 4));
}

// ONLINE = false: two passes (row maxima first).  ONLINE = true: one pass with a LAZY online softmax — the reference
// maximum of a row is only raised (and O, l rescaled through TMEM) when the block's maximum exceeds it by more than 8
// (log2 units), so P stays <= 256 and the rescale is rare; the log-sum-exp written at the end is exact either way.
template <bool ONLINE>
__global__ void __launch_bounds__(THREADS, 1)
fa_tc_fwd_kernel(const __grid_constant__ CUtensorMap map, const Params p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_empty[2], p_full,
      p_empty, o_full;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int nb = p.T / BN;
  const int row0 = n * p.T;                       // first token row of this image in the [N*T, 3C] matrix
  const int cq = h * 3 * D, ck = cq + D, cv = cq + 2 * D;

  if (tid == 0) {
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1);
      mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1);
      mbar_init(smem_u32(&v_empty[s]), 1);
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 128);
    }
    mbar_init(smem_u32(&p_full), 128);
    mbar_init(smem_u32(&p_empty), 1);
    mbar_init(smem_u32(&o_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS[2] = {tmem, tmem + 128u};
  const uint32_t tO = tmem + 256u;

  if (warp == 5) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
      mbar_expect_tx(smem_u32(&q_full), TILE);
      tma_load_2d(base + OFF_Q, &map, smem_u32(&q_full), cq, row0 + qb * BM);
      const int first2 = ONLINE ? 0 : nb;           // first ring index of the pass that also needs V
      for (int i = 0; i < first2 + nb; ++i) {       // K ring: pass 1 (i < nb) and pass 2 walk the same key blocks
        const int s = i & 1, u = i >> 1, j = i < first2 ? i : i - first2;
        if (u > 0) mbar_wait(smem_u32(&k_empty[s]), (u - 1) & 1);
        mbar_expect_tx(smem_u32(&k_full[s]), TILE);
        tma_load_2d(base + OFF_K + s * TILE, &map, smem_u32(&k_full[s]), ck, row0 + j * BN);
        if (i >= first2) {                           // the P V pass also needs V_j
          const int sv = j & 1, uv = j >> 1;
          if (uv > 0) mbar_wait(smem_u32(&v_empty[sv]), (uv - 1) & 1);
          mbar_expect_tx(smem_u32(&v_full[sv]), TILE);
          tma_load_2d(base + OFF_V + sv * TILE, &map, smem_u32(&v_full[sv]), cv, row0 + j * BN);
        }
      }
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptors: D = f32, A = B = bf16.  S: both K-major, N = 128.  O: B (= V) MN-major, N = 64.
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | (static_cast<uint32_t>(D >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
      const uint64_t dq = make_desc_sw128(base + OFF_Q), dp = make_desc_sw128(base + OFF_P);
      auto qk = [&](int i) {                          // S[i & 1] = Q K_i^T
        const int s = i & 1, u = i >> 1;
        mbar_wait(smem_u32(&k_full[s]), u & 1);
        if (u > 0) mbar_wait(smem_u32(&s_empty[s]), (u - 1) & 1);
        fence_after();
        const uint64_t dk = make_desc_sw128(base + OFF_K + s * TILE);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks)
          umma_bf16(tS[s], dq + static_cast<uint64_t>(ks * 2), dk + static_cast<uint64_t>(ks * 2), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(smem_u32(&k_empty[s]));
        umma_commit(smem_u32(&s_full[s]));
      };
      mbar_wait(smem_u32(&q_full), 0);
      const int first2 = ONLINE ? 0 : nb;
      for (int i = 0; i < first2; ++i) qk(i);         // pass 1: scores for the row maxima
      qk(first2);                                     // block 0 of the P V pass
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) qk(first2 + j + 1);           // next block's scores run under this block's exponentials
        const int sv = j & 1, uv = j >> 1;
        mbar_wait(smem_u32(&p_full), j & 1);
        mbar_wait(smem_u32(&v_full[sv]), uv & 1);
        fence_after();
        const uint64_t dv = make_desc_sw128(base + OFF_V + sv * TILE);
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          // A = P: 16 keys = 32 B inside a 64-key atom (+2), second atom 16 KiB further.
          // B = V, MN-major: 16 keys = two 8-row groups = 2048 B further per step.
          const uint64_t pa = dp + static_cast<uint64_t>((ks >> 2) * (TILE >> 4) + (ks & 3) * 2);
          const uint64_t vb = dv + static_cast<uint64_t>(ks * (2048 >> 4));
          umma_bf16(tO, pa, vb, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&p_empty));
        umma_commit(smem_u32(&v_empty[sv]));
      }
      umma_commit(smem_u32(&o_full));
    }
  } else {
    // ===== softmax warps: thread = query row =====
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    float mraw = -INFINITY;
    const int first2 = ONLINE ? 0 : nb;
    for (int i = 0; i < first2; ++i) {                // pass 1
      const int s = i & 1, u = i >> 1;
      mbar_wait(smem_u32(&s_full[s]), u & 1);
      fence_after();
      {
        // all four TMEM loads in flight before the single wait: one exposed TMEM round trip per block, not four
        // (one warp per scheduler: nothing else hides it)
        float v[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS[s] + lane_addr + static_cast<uint32_t>(c * 32), v + c * 32);
        tmem_ld_wait();
        fence_before();
        mbar_arrive(smem_u32(&s_empty[s]));           // the values are in registers: the buffer is free
        float m4[4] = {mraw, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int k = 0; k < 128; k += 4) {
          m4[0] = fmaxf(m4[0], v[k]); m4[1] = fmaxf(m4[1], v[k + 1]);
          m4[2] = fmaxf(m4[2], v[k + 2]); m4[3] = fmaxf(m4[3], v[k + 3]);
        }
        mraw = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
    }
    float m = ONLINE ? -INFINITY : mraw * p.scale_log2;   // scale > 0: max commutes with it
    float l = 0.f;
    const uint32_t pbase = base + OFF_P;
    for (int j = 0; j < nb; ++j) {                    // pass 2
      const int i = first2 + j, s = i & 1, u = i >> 1;
      mbar_wait(smem_u32(&s_full[s]), u & 1);
      fence_after();
      uint32_t w[64];
      bool waited_p = false;
      {
        float v[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS[s] + lane_addr + static_cast<uint32_t>(c * 32), v + c * 32);
        tmem_ld_wait();
        fence_before();
        mbar_arrive(smem_u32(&s_empty[s]));           // S is in registers: the next QK^T may overwrite the buffer
        if (ONLINE) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int k = 0; k < 128; k += 4) {
            m4[0] = fmaxf(m4[0], v[k]); m4[1] = fmaxf(m4[1], v[k + 1]);
            m4[2] = fmaxf(m4[2], v[k + 2]); m4[3] = fmaxf(m4[3], v[k + 3]);
          }
          const float mloc = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
          const bool need = mloc > m + 8.0f;          // always true for block 0 (m = -inf)
          if (__any_sync(0xffffffffu, need)) {        // tcgen05.ld / st are warp-collective: the whole warp goes
            const float m_new = need ? mloc : m;
            const float alpha = j == 0 ? 0.f : exp2f(m - m_new);      // 1 for the rows that keep their reference
            l *= alpha;
            m = m_new;
            if (j > 0) {
              mbar_wait(smem_u32(&p_empty), (j - 1) & 1);             // O is complete up to block j-1
              waited_p = true;
              fence_after();
#pragma unroll 1
              for (int c = 0; c < 2; ++c) {
                float o[32];
                tmem_ld32(tO + lane_addr + static_cast<uint32_t>(c * 32), o);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) o[k] *= alpha;
                tmem_st32(tO + lane_addr + static_cast<uint32_t>(c * 32), o);
              }
              tmem_st_wait();
              fence_before();
            }
          }
        }
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const float e0 = exp2f(fmaf(v[2 * k], p.scale_log2, -m)), e1 = exp2f(fmaf(v[2 * k + 1], p.scale_log2, -m));
          l4[k & 3] += e0 + e1;
          w[k] = pack_bf16x2(e0, e1);
        }
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      }
      if (j > 0 && !waited_p) mbar_wait(smem_u32(&p_empty), (j - 1) & 1);      // the previous PV product has consumed P
#pragma unroll
      for (int q8 = 0; q8 < 16; ++q8)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pbase + p_chunk_off(r, q8)), "r"(w[4 * q8]),
                     "r"(w[4 * q8 + 1]), "r"(w[4 * q8 + 2]), "r"(w[4 * q8 + 3])
                     : "memory");
      fence_async_smem();
      mbar_arrive(smem_u32(&p_full));
    }
    // epilogue: O / l -> bf16, log-sum-exp in the log2 domain
    mbar_wait(smem_u32(&o_full), 0);
    fence_after();
    const float inv = 1.0f / l;
    const size_t C = static_cast<size_t>(p.heads) * D;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(row0) + qb * BM + r) * C + static_cast<size_t>(h) * D;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(tO + lane_addr + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) {
        uint4 o4;
        o4.x = pack_bf16x2(v[8 * q8] * inv, v[8 * q8 + 1] * inv);
        o4.y = pack_bf16x2(v[8 * q8 + 2] * inv, v[8 * q8 + 3] * inv);
        o4.z = pack_bf16x2(v[8 * q8 + 4] * inv, v[8 * q8 + 5] * inv);
        o4.w = pack_bf16x2(v[8 * q8 + 6] * inv, v[8 * q8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c * 32 + q8 * 8) = o4;
      }
    }
    p.lse[(static_cast<size_t>(n) * p.heads + h) * p.T + qb * BM + r] = m + log2f(l);
    fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace fatc

int attention_tc_init() {
  ISB_CUDA(cudaFuncSetAttribute(fatc::fa_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fatc::SMEM_BYTES));
  ISB_CUDA(cudaFuncSetAttribute(fatc::fa_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fatc::SMEM_BYTES));
  return ISB_OK;
}

// ISB_FA_TC: 0 = mma.sync kernel everywhere (default until the tcgen05 variants are measured faster),
// 1 = tcgen05 two-pass, 2 = tcgen05 lazy-online
static int fa_tc_mode() {
  static const int mode = [] {
    const char* e = getenv("ISB_FA_TC");
    return e ? atoi(e) : 0;
  }();
  return mode;
}
bool attention_tc_usable(int T, int ch) { return fa_tc_mode() != 0 && ch == fatc::D && T % fatc::BN == 0 && T >= fatc::BN; }

int attention_tc_forward(const void* qkv, int N, int T, int heads, void* out, float* lse, float scale_log2, cudaStream_t st) {
  CUtensorMap map;
  const cuuint64_t C3 = static_cast<cuuint64_t>(3) * heads * fatc::D;
  cuuint64_t gdim[2] = {C3, static_cast<cuuint64_t>(N) * T};
  cuuint64_t gstr[1] = {C3 * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_tensormap_encode()(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_tc: cuTensorMapEncodeTiled failed: %d", static_cast<int>(r));
    return ISB_ERR_CUDA;
  }
  fatc::Params p{static_cast<__nv_bfloat16*>(out), lse, T, heads, scale_log2};
  if (fa_tc_mode() == 1)
    ISB_CUDA(isb::launch(fatc::fa_tc_fwd_kernel<false>, dim3(T / fatc::BM, heads, N), dim3(fatc::THREADS), fatc::SMEM_BYTES, st, map, p));
  else
    ISB_CUDA(isb::launch(fatc::fa_tc_fwd_kernel<true>, dim3(T / fatc::BM, heads, N), dim3(fatc::THREADS), fatc::SMEM_BYTES, st, map, p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // namespace isb
