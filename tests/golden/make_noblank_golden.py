"""Generates tests/golden/noblank_*.npz by running the REFERENCE's own numba code for the blank-free CTC variant
(/root/reference/pytorch_end2end/functions/ctc_without_blank.py: _ctc_without_blank_3d_loss :91-117 over
_ctc_without_blank_loss :13-88, and the module modules/ctc_without_blank.py) in the authoring container.

    python tests/golden/make_noblank_golden.py
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
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


for pkg in ("pytorch_end2end", "pytorch_end2end.functions"):
    sys.modules.setdefault(pkg, type(sys)(pkg))
load("pytorch_end2end.functions.utils", os.path.join(REF, "pytorch_end2end", "functions", "utils.py"))
ref = load("pytorch_end2end.functions.ctc_without_blank", os.path.join(REF, "pytorch_end2end", "functions", "ctc_without_blank.py"))
mod = load("ref_noblank_module", os.path.join(REF, "pytorch_end2end", "modules", "ctc_without_blank.py"))


def draw(B, T, V, Lmin, Lmax, seed, scale=1.0, lo=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.log_softmax(torch.randn(B, T, V, generator=g) * scale, 2)
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    tg = torch.randint(lo, V, (B, Lmax), generator=g)
    ll = torch.randint(max(Lmax + 2, 3 * T // 4), T + 1, (B,), generator=g)
    return x, tg, ll, tl


def save(name, x, tg, ll, tl, space):
    losses, grads = ref._ctc_without_blank_3d_loss(x.numpy(), tg.numpy(), ll.numpy(), tl.numpy(), space)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), log_probs=x.numpy(), targets=tg.numpy(), logits_lengths=ll.numpy(),
                        targets_lengths=tl.numpy(), space_idx=np.array(space), losses=losses, grads=grads)
    print(name, tuple(x.shape), "space", space, "losses", np.round(losses[:4], 4))


save("noblank_c1", *draw(4, 50, 28, 10, 29, 0), -1)
save("noblank_c1_space", *draw(4, 50, 28, 10, 29, 1), 0)
save("noblank_c2_b4", *draw(4, 400, 29, 100, 200, 2), -1)
save("noblank_c2_b4_space", *draw(4, 400, 29, 100, 200, 3, scale=4.0), 28)
# edge cases: no targets, the single target is the space, tight (T == L / L + 2 needs at least L frames), T == 1, infeasible
x, tg, ll, tl = draw(7, 16, 5, 1, 6, 9, lo=0)
tl[:] = torch.tensor([0, 1, 6, 1, 3, 6, 2])
ll[:] = torch.tensor([16, 9, 6, 1, 16, 4, 16])
tg[1, 0] = 2                                            # the single target IS the space (space_idx = 2 below)
save("noblank_edge_space", x, tg, ll, tl, 2)
tl2 = tl.clone(); tl2[0] = 1                            # space_idx = -1 with no targets indexes column -1 in the reference: keep L >= 1
save("noblank_edge", x, tg, ll, tl2, -1)

# the module (modules/ctc_without_blank.py:7-35) with autograd through its LogSoftmax
g = torch.Generator().manual_seed(5)
logits = torch.randn(3, 30, 7, generator=g)
tgm = torch.randint(0, 7, (3, 8), generator=g)
llm, tlm = torch.tensor([30, 25, 28]), torch.tensor([8, 5, 7])
res = {}
for space in (-1, 3):
    for reduce in (True, False):
        leaf = logits.clone().requires_grad_()
        loss = mod.CTCWithoutBlankLoss(reduce=reduce, space_idx=space)(leaf, tgm, llm, tlm)
        loss.sum().backward()
        res["loss_s%d_r%d" % (space, reduce)] = loss.detach().numpy()
        res["grad_s%d_r%d" % (space, reduce)] = leaf.grad.numpy()
np.savez_compressed(os.path.join(HERE, "noblank_module.npz"), logits=logits.numpy(), targets=tgm.numpy(), logits_lengths=llm.numpy(),
                    targets_lengths=tlm.numpy(), **res)
print("noblank_module", {k: np.round(v, 4).tolist() for k, v in res.items() if k.startswith("loss")})
