"""TEST INFRASTRUCTURE ONLY.

`oracle/` is a CPU restatement (plain PyTorch fp32) of the reference's UNet hot
path (alibaba/VideoMV `tools/modules/unet/{unet_t2v,unet_i2vgen,util}.py`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product (`videomv_b200/`) never does.

Parity pin: the reference ships no golden vectors, so the
restatement is pinned against the *reference itself* imported in the build
container (`oracle/ref_import.py`) by `oracle/gen_golden.py`, which also writes
the committed fixtures in `tests/golden/`.
"""
