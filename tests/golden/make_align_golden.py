"""Generates tests/golden/align_*.npz by running the REFERENCE's own numba alignment code in the authoring container.

Loads /root/reference/pytorch_end2end/utils/alignment.py by path (``get_alignment_3d`` :109-138 over
``_get_alignment_ctc_1d`` :50-106 and ``_get_alignment_asg_1d`` :9-47) and the consumer module
``AlignedTargetsLoss`` (modules/alignment_loss.py:7-33).  /root/reference does not exist on the GPU box, so the
vectors are committed; this script is the record of how they were made.

    python tests/golden/make_align_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("E2E_REFERENCE_ROOT", "/root/reference")


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref = load("ref_alignment", os.path.join(REF, "pytorch_end2end", "utils", "alignment.py"))


def draw(B, T, V, Lmin, Lmax, seed, scale=1.0, full=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.log_softmax(torch.randn(B, T, V, generator=g) * scale, 2)
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    tg = torch.randint(1, V, (B, Lmax), generator=g)
    ll = torch.full((B,), T, dtype=torch.int64) if full else torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    return x, tg, ll, tl


def save(name, x, tg, ll, tl, is_ctc):
    out = ref.get_alignment_3d(x, tg, ll, tl, is_ctc=is_ctc)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), log_probs=x.numpy(), targets=tg.numpy(), logits_lengths=ll.numpy(),
                        targets_lengths=tl.numpy(), is_ctc=np.array(int(is_ctc)), aligned=out.numpy())
    print(name, tuple(x.shape), "->", out.shape, "labels used", int((out > 0).sum()), "ignored", int((out == -100).sum()))


# the BASELINE shapes at oracle-friendly batch sizes
save("align_c1", *draw(4, 50, 28, 10, 29, 0, full=True), True)
save("align_c2_b8", *draw(8, 400, 29, 100, 200, 1), True)
save("align_c2_b4_peaky", *draw(4, 400, 29, 100, 200, 11, scale=8.0), True)
save("align_c4_b2", *draw(2, 250, 1024, 40, 80, 3), True)
save("align_c5_b2", *draw(2, 1600, 29, 300, 600, 4), True)
save("align_asg_c1", *draw(4, 50, 28, 10, 29, 5, full=True), False)
save("align_asg_c2_b4", *draw(4, 400, 29, 100, 200, 6), False)
# edge cases: no targets, one frame, tight (T == L + repeats), repeats, exact ties (uniform log-probs), short slices
x, tg, ll, tl = draw(8, 24, 6, 1, 8, 21)
tl[:] = torch.tensor([0, 1, 8, 3, 5, 2, 8, 4])
ll[:] = torch.tensor([24, 1, 24, 7, 24, 2, 17, 24])
tg[2] = torch.tensor([1, 1, 2, 2, 3, 3, 4, 4])          # repeats
tg[6] = torch.tensor([5, 5, 5, 5, 5, 5, 5, 5])          # all repeats: needs 15 frames, has 17
tg[3, :3] = torch.tensor([2, 2, 2])                     # needs 5 frames, has 7
x[4] = torch.log(torch.full((24, 6), 1.0 / 6))          # exact ties everywhere: the comparison order decides
x[7, :, 0] = x[7, :, 1]                                 # blank ties with label 1
save("align_edge", x, tg, ll, tl, True)
x2, tg2, ll2, tl2 = draw(6, 20, 5, 1, 6, 22)
tl2[:] = torch.tensor([1, 6, 3, 2, 6, 4])
ll2[:] = torch.tensor([1, 6, 20, 20, 20, 9])
x2[4] = torch.log(torch.full((20, 5), 0.2))
save("align_asg_edge", x2, tg2, ll2, tl2, False)

# the consumer module (alignment_loss.py:7-33) on one small batch, both flags
sys.modules["pytorch_end2end.utils.alignment"] = ref
sys.modules.setdefault("pytorch_end2end", type(sys)("pytorch_end2end"))
sys.modules.setdefault("pytorch_end2end.utils", type(sys)("pytorch_end2end.utils"))
mod = load("ref_alignment_loss", os.path.join(REF, "pytorch_end2end", "modules", "alignment_loss.py"))
x, tg, ll, tl = draw(5, 40, 12, 3, 10, 31)
res = {}
for is_ctc in (True, False):
    for ib in (False, True):
        res["loss_ctc%d_ib%d" % (is_ctc, ib)] = mod.AlignedTargetsLoss(is_ctc, ignore_blank=ib)(x, tg, ll, tl).numpy()
np.savez_compressed(os.path.join(HERE, "align_loss_module.npz"), log_probs=x.numpy(), targets=tg.numpy(), logits_lengths=ll.numpy(),
                    targets_lengths=tl.numpy(), **res)
print("align_loss_module", {k: v.round(4).tolist() for k, v in res.items()})
