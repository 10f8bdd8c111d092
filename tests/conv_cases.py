"""Shared convolution test cases: (name, N, H, W, Cin, Cout, ksize, Cin2, bias, residual, accumulate, out_bf16, tune)."""
CASES = [
    ("k3_16x16_64_64", 1, 16, 16, 64, 64, 3, 0, True, False, False, False, None),
    ("k3_32x32_128_256_res", 1, 32, 32, 128, 256, 3, 0, True, True, False, False, None),
    ("k3_8x8_n2_128_128", 2, 8, 8, 128, 128, 3, 0, True, False, False, False, None),
    ("k3_8x8_n1_256_256_splitk", 1, 8, 8, 256, 256, 3, 0, True, True, False, False, None),
    ("k3_8x8_n3_64_64", 3, 8, 8, 64, 64, 3, 0, False, False, False, False, None),
    ("k3_64x64_64_192", 1, 64, 64, 64, 192, 3, 0, True, False, False, False, None),
    ("k1_32x32_256_768", 1, 32, 32, 256, 768, 1, 0, True, False, False, False, None),
    ("k3_16x16_128_128_skip192", 1, 16, 16, 128, 128, 3, 192, True, False, False, False, None),
    ("k3_128x128_64_64", 1, 128, 128, 64, 64, 3, 0, True, False, False, False, None),
    ("k3_32x32_64_128_bf16out", 1, 32, 32, 64, 128, 3, 0, True, False, False, True, None),
    ("k3_32x32_64_128_acc", 1, 32, 32, 64, 128, 3, 0, True, True, True, False, None),
    ("k3_64x64_128_256_bn256", 1, 64, 64, 128, 256, 3, 0, True, False, False, False, {"block_n": 256, "stages": 4}),
    ("k3_32x32_128_128_split3", 1, 32, 32, 128, 128, 3, 0, True, True, False, False, {"split_k": 3}),
    ("k3_16x16_128_64_bn64_st2", 1, 16, 16, 128, 64, 3, 0, True, False, False, False, {"block_n": 64, "stages": 2}),
    ("k1_16x16_n2_192_256", 2, 16, 16, 192, 256, 1, 0, False, True, False, False, None),
    # cluster split-K (DSMEM fold): 2, 4 and 8 CTAs per output tile, 128- and 256-wide tiles, bf16 out
    ("k3_32x32_128_128_cl2", 1, 32, 32, 128, 128, 3, 0, True, True, False, False, {"split_k": 2}),
    ("k3_16x16_256_256_cl4", 1, 16, 16, 256, 256, 3, 0, True, False, True, False, {"split_k": 4, "block_n": 128}),
    ("k3_16x16_256_256_cl8_bn256", 1, 16, 16, 256, 256, 3, 0, True, True, False, False, {"split_k": 8, "block_n": 256}),
    # CTA pairs (cta_group::2): 256-row MMA, B tile split across the two CTAs
    ("k3_32x32_128_256_pair", 1, 32, 32, 128, 256, 3, 0, True, True, False, False, {"two_cta": 1, "block_n": 256, "split_k": 1}),
    ("k3_64x64_64_128_pair_bn128", 1, 64, 64, 64, 128, 3, 0, True, False, False, True, {"two_cta": 1, "block_n": 128, "split_k": 1}),
    ("k3_16x16_n2_128_512_pair_skip", 2, 16, 16, 128, 512, 3, 64, True, False, True, False, {"two_cta": 1, "block_n": 256, "split_k": 1}),
    ("k1_128x128_64_256_pair", 1, 128, 128, 64, 256, 1, 0, False, True, False, False, {"two_cta": 1, "block_n": 256, "split_k": 1}),
    ("k1_8x8_1024_1024_auto", 1, 8, 8, 1024, 1024, 1, 0, True, True, False, True, None),
    ("k3_8x8_1024_1024_auto", 1, 8, 8, 1024, 1024, 3, 0, True, False, False, False, None),
]
