"""Differential fuzz of the beam-search kernel against the C restatement: python tools/beam_fuzz.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200.engine import CTCBeamEngine
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = torch.Generator().manual_seed(seed)
t0 = time.time(); n = 0; bad = 0
def ri(a, b): return int(torch.randint(a, b + 1, (1,), generator=g))
while time.time() - t0 < budget:
    V = [2, 3, 4, 7, 29, 64, 150, 260][ri(0, 7)]
    beam = [1, 2, 5, 16, 50, 100, 256][ri(0, 6)]
    if V * beam > 40000: continue
    T = ri(1, 120)
    B = [1, 3, 8, 160][ri(0, 3)] if V * beam * T < 200000 else ri(1, 4)
    scale = [0.2, 1.0, 3.0, 8.0][ri(0, 3)]
    blank = ri(0, V - 1)
    space = ri(-1, V - 1)
    if space == blank: space = -1
    wip = [0.0, 0.0, 0.5, -0.4, 2.0][ri(0, 4)]
    lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * scale, 2)
    if ri(0, 4) == 0: lp[:, :, ri(0, V - 1)] = float("-inf")       # a masked symbol
    ll = torch.randint(0, T + 1, (B,), generator=g)
    labels = None
    if space >= 0:
        labels = ["x"] * V; labels[space] = " "
    eng = CTCBeamEngine(blank, beam, labels, wip)
    dec, n_out = eng.decode(lp.cuda(), ll)
    port = oracle.beam_decode(lp, ll, blank_idx=blank, beam_width=beam, labels=labels, after_logsoftmax=True, wip=wip, prefer="port", return_ties=True)
    ok = torch.equal(n_out, port[1]) and all(dec[i, :int(n_out[i])].tolist() == port[0][i, :int(n_out[i])].tolist() for i in range(B)) and torch.equal(eng.last_ties, port[3])
    n += 1
    if not ok:
        bad += 1
        print("MISMATCH V=%d beam=%d T=%d B=%d scale=%g blank=%d space=%d wip=%g" % (V, beam, T, B, scale, blank, space, wip), flush=True)
print("beam fuzz: %d cases, %d mismatches, %.0f s" % (n, bad, time.time() - t0))
