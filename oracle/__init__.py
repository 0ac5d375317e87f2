"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Long-Tail-GAN hot path (ash-shar/Long-Tail-GAN).

Nothing under oracle/ is imported by the product package (long-tail-gan_b200/). Only tests/, the `smoke()`
check in __graft_entry__.py and the cpu_baseline / `--impl reference` legs of bench.py use it, and there only as
the checker or as the timed CPU baseline, never as a fallback for the CUDA path.
"""
