"""B200-native stand-in for the reference's ``rocwmma_fattn`` package
(/root/reference/rocwmma_fattn/): import ``FlashAttentionFunction`` from ``rocwmma_fattn.FlashAttn``
exactly as the reference's scripts do (bench_with_sdpa.py:60, precision_test.py:41)."""
