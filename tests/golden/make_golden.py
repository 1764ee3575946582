"""Generates tests/golden/*.npz by running the REFERENCE ITSELF in the authoring container.

Imports the reference's own Python package from /root/reference (pytorch_end2end/modules/ctc_loss.py,
functions/forward_backward.py, decoders/ctc_decoder.py) on top of its own C++ engines compiled
unmodified by oracle/build_ref.py (oracle/_ref/).  /root/reference does not exist on the GPU box,
so the vectors are committed; this script is the record of how they were made.

    python oracle/build_ref.py && python tests/golden/make_golden.py

Inputs are stored exactly (float32 / float64 / raw bf16 bits); outputs are the reference's.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("E2E_REFERENCE_ROOT", "/root/reference")
os.environ.setdefault("OMP_NUM_THREADS", "8")

# the reference package must win over this repo's drop-in package of the same name
sys.path = [REF, os.path.join(ROOT, "oracle", "_ref", "loss"), os.path.join(ROOT, "oracle", "_ref", "decoder")] + \
           [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
import torch  # noqa: E402
import pytorch_end2end as ref  # noqa: E402

assert os.path.abspath(ref.__file__).startswith(REF), ref.__file__
sys.path.insert(0, ROOT)


def t2n(t):
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy()
    return t.detach().numpy()


def run_module(x, tg, ll, tl, **kw):
    """loss (as returned) and leaf gradient of the reference's CTCLoss module."""
    crit = ref.CTCLoss(**kw)
    leaf = x.clone().requires_grad_()
    loss = crit(leaf, tg, ll, tl)
    (loss.sum() if loss.dim() else loss).backward()
    return loss.detach(), leaf.grad.detach()


def engine(blank):
    import cpp_ctc_loss
    return cpp_ctc_loss.CTCLossEngine(blank)


def draw(B, T, V, Lmin, Lmax, seed, full_length=False, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, V, generator=g) * scale
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    tg = torch.randint(1, V, (B, Lmax), generator=g)
    ll = torch.full((B,), T, dtype=torch.int64) if full_length else torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    return x, tg, ll, tl


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (t2n(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("%-14s %7.1f kB" % (name, os.path.getsize(path) / 1e3))


def loss_case(name, x, tg, ll, tl, blank=0, module_modes=()):
    """Engine contract on log-probs + module runs on raw logits for the listed flag sets."""
    lp = torch.log_softmax(x, 2)
    losses, grads = engine(blank).compute(lp, tg, ll, tl)
    out = dict(x=x, targets=tg, logits_lengths=ll, targets_lengths=tl, blank=blank,
               engine_losses=losses, engine_grads=grads)
    for i, kw in enumerate(module_modes):
        xin = x
        if kw.get("after_logsoftmax"):
            xin = lp
        if kw.get("time_major"):
            xin = xin.permute(1, 0, 2).contiguous()
        loss, grad = run_module(xin, tg, ll, tl, blank_idx=blank, **kw)
        out["m%d_flags" % i] = np.array([int(bool(kw.get(k))) for k in ("reduce", "size_average", "after_logsoftmax", "time_major")])
        out["m%d_loss" % i] = loss
        out["m%d_grad" % i] = grad
    save(name, **out)


def main():
    modes = (dict(reduce=True, size_average=True), dict(reduce=True), dict(),
             dict(reduce=True, size_average=True, after_logsoftmax=True),
             dict(reduce=True, time_major=True), dict(after_logsoftmax=True, time_major=True))

    # ---- known-answer vectors of the reference's own tests (tests/test_ctc.py:69-165) ----------
    p2 = [[[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.6, 0.1, 0.1]],
          [[0.6, 0.1, 0.1, 0.1, 0.1], [0.1, 0.1, 0.5, 0.2, 0.1]]]
    tf1 = [[0.0260553, 0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857],
           [0.010436, 0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609],
           [0.0037688, 0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882],
           [0.00331533, 0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545],
           [0.00623107, 0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441]]
    tf2 = [[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
           [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
           [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
           [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
           [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]]
    # (the two-utterance literals are written frame-major in the reference's test and transposed there)
    kats = [  # (name, batch-major input, is it log-probs?, targets, lengths, blank, expected cost)
        ("simple", torch.tensor(p2[:1]), False, [[1, 2]], [2], [2], 0, 2.4628584384918),
        ("medium", torch.tensor(p2).transpose(0, 1).contiguous(), False, [[1, 2], [1, 2]], [2, 2], [2, 2], 0, 6.0165174007416),
        ("empty_label", torch.tensor(p2).transpose(0, 1).contiguous(), False, [[1, 2], [0, 0]], [2, 2], [2, 0], 0, 6.416517496109),
        ("tf_1", torch.log(torch.tensor([tf1])), True, [[1, 2, 3, 2, 1]], [5], [5], 5, 3.34211),
        ("tf_2", torch.log(torch.tensor([tf2])), True, [[0, 1, 1, 0]], [5], [4], 5, 5.42262),
    ]
    out = {}
    for name, x, is_lp, tg, ll, tl, blank, expected in kats:
        tg, ll, tl = torch.IntTensor(tg), torch.IntTensor(ll), torch.IntTensor(tl)
        lp = x if is_lp else torch.log_softmax(x, -1)      # what tests/test_ctc.py:36-39 feeds the module
        loss, grad = run_module(lp.permute(1, 0, 2).contiguous(), tg, ll, tl, reduce=True, size_average=False,
                                after_logsoftmax=True, time_major=True, blank_idx=blank)
        assert abs(loss.item() - expected) < 1e-5, (name, loss.item(), expected)
        out.update({name + "_lp": lp, name + "_targets": tg, name + "_ll": ll, name + "_tl": tl,
                    name + "_blank": blank, name + "_expected": expected, name + "_ref_loss": loss,
                    name + "_ref_grad_tm": grad})
    save("kat_loss", **out)

    # ---- BASELINE configs[0]: README example shape, full size ------------------------------------
    x, tg, ll, tl = draw(4, 50, 28, 10, 29, 0, full_length=True)
    loss_case("c1", x, tg, ll, tl, module_modes=modes)
    # ---- LibriSpeech-shaped rows (configs[1]) at B=4, plus a peaky (x5) variant -----------------
    x, tg, ll, tl = draw(4, 400, 29, 100, 200, 11)
    loss_case("c2_b4", x, tg, ll, tl, module_modes=modes[:1])
    x, tg, ll, tl = draw(4, 400, 29, 100, 200, 12, scale=5.0)
    loss_case("c2_b4_peaky", x, tg, ll, tl, module_modes=modes[:1])
    # ---- subword shape (configs[3]) at B=2 ---------------------------------------------------------
    x, tg, ll, tl = draw(2, 250, 1024, 40, 80, 13)
    loss_case("c4_b2", x[:, :160, :200].contiguous(), tg.clamp(max=199), ll.clamp(max=160), tl, module_modes=modes[:1])
    # ---- edge cases (SURVEY 8a notes): infeasible, L=0, T_i=1, repeats, blank!=0, int32 ----------
    g = torch.Generator().manual_seed(21)
    x = torch.randn(8, 9, 6, generator=g)
    tg = torch.tensor([[1, 1, 2, 0], [1, 2, 3, 4], [2, 2, 2, 2], [4, 0, 0, 0], [1, 2, 1, 0], [3, 3, 1, 0],
                       [5, 4, 5, 4], [1, 0, 0, 0]], dtype=torch.int32)
    tl = torch.tensor([3, 4, 4, 1, 0, 2, 4, 1], dtype=torch.int32)
    ll = torch.tensor([9, 4, 6, 1, 5, 2, 9, 9], dtype=torch.int32)   # rows 2 and 5 are infeasible
    loss_case("edge_blank0", x, tg, ll, tl, blank=0, module_modes=(dict(), dict(after_logsoftmax=True), dict(time_major=True)))
    tg3 = torch.where(tg == 3, torch.tensor(0, dtype=torch.int32), tg)  # blank=3 must not occur as a label
    loss_case("edge_blank3", x, tg3, ll, tl, blank=3, module_modes=(dict(), dict(after_logsoftmax=True)))
    # ---- double precision (the reference's gradcheck fixture shape, tests/test_ctc.py:168-191) ---
    rs = np.random.RandomState(678)
    tl64 = rs.randint(1, 11, size=2)
    ll64 = tl64 + rs.randint(0, 11, size=2)
    x64 = torch.from_numpy(rs.randn(2, 20, 6))
    tg64 = torch.from_numpy((1 + rs.rand(2, tl64.max()) * 5).astype(np.int64))
    loss, grad = run_module(x64, tg64, torch.from_numpy(ll64), torch.from_numpy(tl64), blank_idx=0)
    save("f64", x=x64, targets=tg64, logits_lengths=ll64, targets_lengths=tl64, m0_loss=loss, m0_grad=grad)
    # ---- bf16 (configs[2] shape at B=4): oracle is the reference on logits.float() --
    x, tg, ll, tl = draw(4, 128, 96, 20, 40, 14)
    xb = x.to(torch.bfloat16)
    loss, grad = run_module(xb.float(), tg, ll, tl, reduce=True, size_average=True)
    loss_u, _ = run_module(xb.float(), tg, ll, tl)
    save("bf16_c3_b4", x_bf16_bits=xb, targets=tg, logits_lengths=ll, targets_lengths=tl,
         m0_loss=loss, m0_grad=grad, per_utt_loss=loss_u)

    # ---- greedy decoder: the reference's known answers (tests/test_ctc_decoder.py:44-166) --------
    labels7 = ["'", " ", "a", "b", "c", "d", "_"]
    probs1 = [[0.06390443, 0.21124858, 0.27323887, 0.06870235, 0.0361254, 0.18184413, 0.16493624],
              [0.03309247, 0.22866108, 0.24390638, 0.09699597, 0.31895462, 0.0094893, 0.06890021],
              [0.218104, 0.19992557, 0.18245131, 0.08503348, 0.14903535, 0.08424043, 0.08120984],
              [0.12094152, 0.19162472, 0.01473646, 0.28045061, 0.24246305, 0.05206269, 0.09772094],
              [0.1333387, 0.00550838, 0.00301669, 0.21745861, 0.20803985, 0.41317442, 0.01946335],
              [0.16468227, 0.1980699, 0.1906545, 0.18963251, 0.19860937, 0.04377724, 0.01457421]]
    probs2 = [[0.08034842, 0.22671944, 0.05799633, 0.36814645, 0.11307441, 0.04468023, 0.10903471],
              [0.09742457, 0.12959763, 0.09435383, 0.21889204, 0.15113123, 0.10219457, 0.20640612],
              [0.45033529, 0.09091417, 0.15333208, 0.07939558, 0.08649316, 0.12298585, 0.01654384],
              [0.02512238, 0.22079203, 0.19664364, 0.11906379, 0.07816055, 0.22538587, 0.13483174],
              [0.17928453, 0.06065261, 0.41153005, 0.1172041, 0.11880313, 0.07113197, 0.04139363],
              [0.15882358, 0.1235788, 0.23376776, 0.20510435, 0.00279306, 0.05294827, 0.22298418]]
    gk = [("simple", torch.tensor([[[1., 2, 4, 3, 10], [2, 1, 3, 8, 1]]]), [2], 0, ["_", "a", "b", "c", "d"], ["dc"]),
          ("sm", torch.log(torch.tensor([[[0.7, 0.3], [0.7, 0.3]]])), None, 0, ["_", "a"], [""]),
          ("probs_1", torch.log(torch.tensor([probs1])), None, 6, labels7, ["ac'bdc"]),
          ("probs_2", torch.log(torch.tensor([probs2])), None, 6, labels7, ["b'da"])]
    out = {}
    for name, x, ll, blank, labels, expected in gk:
        dec = ref.CTCDecoder(beam_width=1, blank_idx=blank, labels=labels, time_major=False)
        r = dec.decode(x, torch.LongTensor(ll) if ll else None)
        assert r.decoded_sentences == expected, (name, r.decoded_sentences)
        out.update({name + "_x": x, name + "_ll": np.array(ll if ll else [-1]), name + "_blank": blank,
                    name + "_labels": np.array(labels), name + "_sentences": np.array(expected),
                    name + "_targets": r.decoded_targets, name + "_lengths": r.decoded_targets_lengths})
    save("kat_greedy", **out)
    # ---- greedy on random data with exact ties and NaNs, fp32 and bf16 ---------------------------
    g = torch.Generator().manual_seed(31)
    x = torch.randint(-3, 4, (6, 40, 11), generator=g).float()          # many exact ties
    x[1, 3, 4] = float("nan"); x[1, 3, 7] = float("nan"); x[2, 0, 0] = float("nan"); x[3, 39, 10] = float("nan")
    ll = torch.tensor([40, 33, 1, 40, 17, 25])
    dec = ref.CTCDecoder(beam_width=1, blank_idx=2)
    r = dec.decode(x, ll)
    r_full = dec.decode(x, None)
    xb = (torch.randn(4, 64, 96, generator=g)).to(torch.bfloat16)
    llb = torch.tensor([64, 50, 64, 9])
    rb = ref.CTCDecoder(beam_width=1, blank_idx=0).decode(xb, llb)
    rtm = ref.CTCDecoder(beam_width=1, blank_idx=0, time_major=True).decode(xb.transpose(0, 1).contiguous(), llb)
    assert torch.equal(rb.decoded_targets, rtm.decoded_targets)
    save("greedy_random", x=x, ll=ll, blank=2, targets=r.decoded_targets, lengths=r.decoded_targets_lengths,
         targets_full=r_full.decoded_targets, lengths_full=r_full.decoded_targets_lengths,
         xb_bits=xb, llb=llb, targets_b=rb.decoded_targets, lengths_b=rb.decoded_targets_lengths)


if __name__ == "__main__":
    main()
