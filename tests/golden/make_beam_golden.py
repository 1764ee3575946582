"""Generates tests/golden/beam_*.npz by running the REFERENCE's own decoder in the authoring container.

The reference's C++ (src/decoders/ctc_decoder.cpp, compiled unmodified from /root/reference into oracle/_ref by
oracle/build_ref.py) is driven exactly as its Python wrapper drives it (pytorch_end2end/decoders/ctc_decoder.py:76-115:
``cpp_ctc_decoder.CTCDecoder(blank_idx, beam_width, labels, "", lmwt, wip, oov_penalty, case_sensitive).decode(
logits_=log-probabilities, logits_lengths_=...)``).  /root/reference does not exist on the GPU box, so the vectors are
committed; this script is the record of how they were made.

    python tests/golden/make_beam_golden.py

Every case is also run through the C restatement (oracle.beam_decode(prefer="port")), whose tie counter says whether
any prune of an utterance had EQUAL scores on both sides of the cut (there the reference's pick is libstdc++'s
introselect order); the counter is stored next to the reference's answer.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

mod = oracle._load_ref("cpp_ctc_decoder", "decoder")
assert mod is not None, "build oracle/_ref first (python oracle/build_ref.py)"


def run(name, lp, lengths, blank, beam, labels=None, wip=0.0):
    labels = list(labels or [])
    dec = mod.CTCDecoder(int(blank), int(beam), labels, "", 1.0, float(wip), -1000.0, False)
    tgt, tl, sents = dec.decode(logits_=lp.contiguous(), logits_lengths_=lengths)
    port = oracle.beam_decode(lp, lengths, blank_idx=blank, beam_width=beam, labels=labels, after_logsoftmax=True,
                              wip=wip, prefer="port", return_ties=True)
    B = lp.size(0)
    same = [int(tl[i]) == int(port[1][i]) and tgt[i, :int(tl[i])].tolist() == port[0][i, :int(tl[i])].tolist() for i in range(B)]
    ties = port[3].numpy()
    assert all(s or t > 0 for s, t in zip(same, ties)), (name, same, ties)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), log_probs=lp.numpy(), logits_lengths=lengths.numpy(),
                        blank=np.array(blank), beam=np.array(beam), wip=np.array(wip),
                        labels=np.array("\x1f".join(labels)),
                        targets=tgt.numpy(), targets_lengths=tl.numpy(), sentences=np.array("\x1f".join(sents)), ties=ties)
    print(name, tuple(lp.shape), "beam", beam, "lengths", tl.tolist()[:8], "ties", ties.tolist()[:8], "port==ref", all(same))


def draw(B, T, V, seed, scale=1.0, full=False):
    g = torch.Generator().manual_seed(seed)
    lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * scale, 2)
    ll = torch.full((B,), T, dtype=torch.int64) if full else torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    return lp, ll


# the reference's own known answers (tests/test_ctc_decoder.py:86-166)
probs = {
    "sm": ([[[0.7, 0.3], [0.7, 0.3]]], ["_", "a"], 0),
    "1": ([[[0.06390443, 0.21124858, 0.27323887, 0.06870235, 0.0361254, 0.18184413, 0.16493624],
            [0.03309247, 0.22866108, 0.24390638, 0.09699597, 0.31895462, 0.0094893, 0.06890021],
            [0.218104, 0.19992557, 0.18245131, 0.08503348, 0.14903535, 0.08424043, 0.08120984],
            [0.12094152, 0.19162472, 0.01473646, 0.28045061, 0.24246305, 0.05206269, 0.09772094],
            [0.1333387, 0.00550838, 0.00301669, 0.21745861, 0.20803985, 0.41317442, 0.01946335],
            [0.16468227, 0.1980699, 0.1906545, 0.18963251, 0.19860937, 0.04377724, 0.01457421]]],
          ["'", " ", "a", "b", "c", "d", "_"], 6),
    "2": ([[[0.08034842, 0.22671944, 0.05799633, 0.36814645, 0.11307441, 0.04468023, 0.10903471],
            [0.09742457, 0.12959763, 0.09435383, 0.21889204, 0.15113123, 0.10219457, 0.20640612],
            [0.45033529, 0.09091417, 0.15333208, 0.07939558, 0.08649316, 0.12298585, 0.01654384],
            [0.02512238, 0.22079203, 0.19664364, 0.11906379, 0.07816055, 0.22538587, 0.13483174],
            [0.17928453, 0.06065261, 0.41153005, 0.1172041, 0.11880313, 0.07113197, 0.04139363],
            [0.15882358, 0.1235788, 0.23376776, 0.20510435, 0.00279306, 0.05294827, 0.22298418]]],
          ["'", " ", "a", "b", "c", "d", "_"], 6),
}
for k, (pr, labels, blank) in probs.items():
    lp = torch.log(torch.FloatTensor(pr))
    run("beam_kat_" + k, lp, torch.full((1,), lp.size(1), dtype=torch.int32), blank, 20, labels, 0.0)

# BASELINE shapes at oracle-friendly sizes, the reference's default beam of 100
run("beam_c1", *draw(4, 50, 28, 0, full=True), 0, 100)
run("beam_c2_b4", *draw(4, 400, 29, 1), 0, 100)
run("beam_c2_b4_peaky", *draw(4, 400, 29, 11, scale=5.0), 0, 100)
run("beam_c3_b8", *draw(8, 128, 96, 2), 0, 100)
run("beam_c4_b2", *draw(2, 120, 1024, 3), 0, 100)
# words: a space label and a word insertion penalty (score = log p - words * wip), blank not at 0
labels = [chr(97 + i) for i in range(26)] + [" ", "'", "_"]
run("beam_words", *draw(6, 80, 29, 21, scale=2.0), 28, 50, labels, 0.8)
run("beam_words_neg", *draw(4, 60, 29, 22, scale=1.5), 28, 16, labels, -0.6)
# narrow beams and small alphabets where every prune still has distinct scores on both sides of the cut
run("beam_w2", *draw(8, 60, 12, 23, scale=2.0), 0, 2)
run("beam_w7_v5", *draw(8, 40, 5, 24, scale=1.0), 2, 7)
# short inputs: one frame, two frames, zero frames (the empty prefix comes back as the symbol -1 with length 1)
lp, ll = draw(6, 12, 9, 25)
ll[:] = torch.tensor([1, 2, 0, 3, 12, 5])
run("beam_short", lp, ll, 0, 10)
# blank-dominated rows: the best prefix is empty
g = torch.Generator().manual_seed(26)
x = torch.randn(3, 20, 8, generator=g)
x[:, :, 0] += 9.0
run("beam_empty", torch.log_softmax(x, 2), torch.tensor([20, 15, 20]), 0, 25)
# tiny alphabets with a wide beam: equal (-inf) scores straddle the cut; recorded with their tie counters
run("beam_ties_v2", *draw(4, 40, 2, 27, scale=2.0), 1, 5)
run("beam_ties_v3", *draw(4, 40, 3, 28, scale=3.0), 0, 100)
