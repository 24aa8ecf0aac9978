"""CPU oracle for the DINO training-step hot path of wardvdnb/Multimodal-SSL-AVMNIST.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`multimodal_ssl_avmnist_b200`) may import this
package; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do,
and there only as the checker or the timed CPU baseline — never as the thing shipped.

The reference is pure Python/PyTorch, so the oracle is a restatement in numpy (integer / index work of the
augmentations) and plain torch CPU ops (floating-point network, losses, EMA, Adam).  Every function cites the
reference file:line (relative to /root/reference/AVMNIST_Experiments/) it follows.

Parity pin: the reference has no tests and no golden vectors of its own.  The oracle is
pinned against outputs of the *imported reference itself*, generated in the build container by
`tests/golden/make_golden.py` (committed together with the fixtures it wrote, `tests/golden/*.npz|json`) and
against the known-answer values recorded in SURVEY.md §8c.  The augmentation arithmetic lives in un-vendored
third-party packages (torchvision 0.20.1 / torchaudio 2.5.1 pinned by the reference's requirements.txt:12-16;
0.26.0 / 2.11.0 in this image): the pin is what those packages produce here.
"""
