"""LM-free prefix beam search on the GPU against the reference's decoder.

Reference: src/decoders/ctc_decoder.cpp:153-198 (decode), :353-441 (decode_sentence), :241-309 (get_next_prefix),
pytorch_end2end/decoders/ctc_decoder.py:76-115 (the Python wrapper).  Decoded symbols and lengths are integers: the bar
is bit-exact.  The checker is the compiled reference where it travelled with the repo (oracle/_ref) and the C
restatement (oracle/ctc_oracle.c: ctc_oracle_beam, pinned against the compiled reference by tests/golden/beam_*.npz).

Where EQUAL scores straddle the prune cut the reference's pick is libstdc++'s introselect order; the kernel and the
restatement take the lower position and count the event.  Utterances with a zero tie counter must equal the
reference; all utterances must equal the restatement, tie counters included.
"""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from pytorch_end2end import CTCDecoder
from end2end_b200.engine import CTCBeamEngine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _labels(z):
    s = str(z["labels"])
    return s.split("\x1f") if s else []


def _rows(t, n):
    return [t[i, :int(n[i])].tolist() for i in range(t.size(0))]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "beam_*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("where", ["host", "cuda"])
def test_beam_golden(path, where):
    z = np.load(path, allow_pickle=False)
    lp = torch.from_numpy(z["log_probs"])
    ll = torch.from_numpy(z["logits_lengths"])
    labels = _labels(z)
    dec = CTCDecoder(beam_width=int(z["beam"]), after_logsoftmax=True, blank_idx=int(z["blank"]), labels=labels or None,
                     wip=float(z["wip"]))
    x = lp.cuda() if where == "cuda" else lp
    res = dec.decode(x, ll)
    assert not res.decoded_targets.is_cuda and res.decoded_targets.dtype == torch.int64
    want_t, want_l = torch.from_numpy(z["targets"]), torch.from_numpy(z["targets_lengths"])
    ties = torch.from_numpy(z["ties"])
    got_ties = dec._decoder.last_ties
    assert torch.equal(got_ties, ties), (got_ties, ties)
    free = ties == 0
    assert torch.equal(res.decoded_targets_lengths[free], want_l[free])
    got, want = _rows(res.decoded_targets, res.decoded_targets_lengths), _rows(want_t, want_l)
    for i in range(lp.size(0)):
        if free[i]:
            assert got[i] == want[i], i
    if bool(free.all()):
        assert res.decoded_targets.shape == want_t.shape       # [B, longest result], as ctc_decoder.cpp:186-196
        assert torch.equal(res.decoded_targets, want_t)
        sents = str(z["sentences"]).split("\x1f")
        if labels and all(v >= 0 for r in want for v in r):
            assert res.decoded_sentences == sents
    # every utterance, ties or not, equals the C restatement (same tie rule)
    port = oracle.beam_decode(lp, ll, blank_idx=int(z["blank"]), beam_width=int(z["beam"]), labels=labels,
                              after_logsoftmax=True, wip=float(z["wip"]), prefer="port")
    assert got == _rows(port[0], port[1])


def _check_vs_oracle(x, ll, blank, beam, labels=None, wip=0.0, after_logsoftmax=True, time_major=False, dec_input=None):
    dec = CTCDecoder(beam_width=beam, after_logsoftmax=after_logsoftmax, blank_idx=blank, labels=labels, wip=wip,
                     time_major=time_major)
    res = dec.decode(x if dec_input is None else dec_input, ll)
    port = oracle.beam_decode(x, ll, blank_idx=blank,
                              beam_width=beam, labels=labels, after_logsoftmax=after_logsoftmax, time_major=time_major,
                              wip=wip, prefer="port", return_ties=True)
    got = _rows(res.decoded_targets, res.decoded_targets_lengths)
    assert got == _rows(port[0], port[1])
    if after_logsoftmax:   # identical log-probabilities on both sides: the tie counters agree too (the fused log-softmax
        assert torch.equal(dec._decoder.last_ties, port[3])   # differs from torch's CPU one in the last ulp)
    if oracle._load_ref("cpp_ctc_decoder", "decoder") is not None and not (labels and any(-1 in r for r in got)):
        ref = oracle.beam_decode(x, ll, blank_idx=blank, beam_width=beam, labels=labels, after_logsoftmax=after_logsoftmax,
                                 time_major=time_major, wip=wip, prefer="reference")
        want = _rows(ref[0], ref[1])
        for i, t in enumerate(port[3].tolist()):
            if t == 0:
                assert got[i] == want[i], i
    return res


def test_beam_random_sweep():
    """Alphabets, beam widths, lengths and peakiness, with and without words: every utterance equals the restatement;
    utterances without a tie equal the compiled reference."""
    g = torch.Generator().manual_seed(7)
    n = 0
    for V in (2, 3, 5, 8, 29, 64, 200):
        for beam in (2, 3, 5, 20, 100, 256):
            for T, scale in ((1, 1.0), (2, 0.3), (7, 4.0), (33, 1.0), (70, 2.5)):
                B = 3
                lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * scale, 2)
                ll = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
                labels, wip = None, 0.0
                blank = 0 if V % 2 else V - 1
                if V >= 5 and beam % 2 == 1:
                    labels = [chr(33 + i) for i in range(V)]
                    labels[1 if blank == 0 else 0] = " "
                    wip = 0.7
                _check_vs_oracle(lp, ll, blank, beam, labels, wip)
                n += 1
    assert n == 210


def test_beam_fused_log_softmax_dtypes_and_layouts():
    """Raw logits (the F.log_softmax of decoders/ctc_decoder.py:95-97 fused), bf16 / f16 / f64 inputs, time-major views,
    int32 / missing lengths, CUDA and CPU tensors."""
    g = torch.Generator().manual_seed(8)
    x = torch.randn(5, 90, 29, generator=g) * 2.0
    ll = torch.randint(60, 91, (5,), generator=g)
    _check_vs_oracle(x, ll, 0, 100, after_logsoftmax=False)
    _check_vs_oracle(x, ll.int(), 0, 100, after_logsoftmax=False, dec_input=x.cuda())
    _check_vs_oracle(x, None, 0, 30, after_logsoftmax=False)
    _check_vs_oracle(x.double(), ll, 3, 40, after_logsoftmax=False)
    _check_vs_oracle(x.bfloat16(), ll, 0, 100, after_logsoftmax=False)
    _check_vs_oracle(x.half(), ll, 0, 100, after_logsoftmax=False)
    _check_vs_oracle(torch.log_softmax(x, 2).bfloat16(), ll, 0, 50, after_logsoftmax=True)
    xt = x.transpose(0, 1).contiguous()                     # [T, B, V]
    _check_vs_oracle(xt, ll, 0, 100, after_logsoftmax=False, time_major=True)
    _check_vs_oracle(xt, ll, 0, 100, after_logsoftmax=False, time_major=True, dec_input=xt.cuda())
    big = torch.randn(2, 40, 1024, generator=g) * 3.0
    _check_vs_oracle(big, torch.tensor([40, 31]), 0, 100, after_logsoftmax=False)


def test_beam_large_batch_variant():
    """Batches beyond one CTA per SM run the 256-thread kernel (several CTAs per SM): same results."""
    g = torch.Generator().manual_seed(10)
    for V, beam, T, B in ((12, 16, 40, 170), (29, 100, 30, 150), (300, 64, 12, 160)):
        lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * 2.0, 2)
        ll = torch.randint(T // 2, T + 1, (B,), generator=g)
        dec = CTCDecoder(beam_width=beam, after_logsoftmax=True, blank_idx=0, wip=0.3)   # (the wrapper's default wip is 1.0)
        res = dec.decode(lp.cuda(), ll)
        port = oracle.beam_decode(lp, ll, beam_width=beam, after_logsoftmax=True, wip=0.3, prefer="port", return_ties=True)
        assert _rows(res.decoded_targets, res.decoded_targets_lengths) == _rows(port[0], port[1])
        assert torch.equal(dec._decoder.last_ties, port[3])


def test_beam_full_size_c2():
    """BASELINE config 2 at its full batch (B=64, T=400, V=29), the reference's default beam of 100."""
    x, _, ll, _ = oracle.make_inputs(*oracle.CONFIGS["c2"][:6])
    lp = torch.log_softmax(x, 2)
    res = _check_vs_oracle(lp, ll, 0, 100)
    assert int(res.decoded_targets_lengths.min()) > 100


def test_beam_device_api_and_limits():
    g = torch.Generator().manual_seed(9)
    lp = torch.log_softmax(torch.randn(4, 50, 29, generator=g), 2).cuda()
    eng = CTCBeamEngine(0, 100)
    dec, n, ties = eng.decode_device(lp, None)
    assert dec.is_cuda and dec.shape == (4, 50) and n.shape == (4,) and ties.shape == (4,)
    port = oracle.beam_decode(lp.cpu(), None, beam_width=100, after_logsoftmax=True, prefer="port")
    assert _rows(dec.cpu(), n.cpu()) == _rows(port[0], port[1])
    assert bool((dec.cpu()[torch.arange(50)[None, :] >= n.cpu()[:, None]] == 0).all())   # zero padding
    with pytest.raises(Exception, match="limits"):
        CTCBeamEngine(0, 257).decode_device(lp, None)
    with pytest.raises(NotImplementedError):
        CTCDecoder(beam_width=10, lm_path="/nonexistent.arpa").decode(lp)
    # beam_width == 1 is greedy decoding in the reference's wrapper (decoders/ctc_decoder.py:92-93)
    a = CTCDecoder(beam_width=1).decode(lp)
    b = CTCDecoder(beam_width=1).decode_greedy(lp)
    assert torch.equal(a.decoded_targets, b.decoded_targets)
