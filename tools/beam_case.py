"""Stress one beam-search configuration repeatedly against the C restatement (race hunting): python tools/beam_case.py [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from end2end_b200.engine import CTCBeamEngine
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
g = torch.Generator().manual_seed(10)
for V, beam, T, B in ((12, 16, 40, 170), (29, 100, 30, 150), (12, 16, 40, 100)):
    lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * 2.0, 2)
    ll = torch.randint(T // 2, T + 1, (B,), generator=g)
    port = oracle.beam_decode(lp, ll, beam_width=beam, after_logsoftmax=True, prefer="port", return_ties=True)
    want = [port[0][i, :int(port[1][i])].tolist() for i in range(B)]
    eng = CTCBeamEngine(0, beam)
    x = lp.cuda(); l = ll.cuda()
    nbad = 0; seen = {}
    junk = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for rep in range(reps):
        if rep % 3 == 0: junk.random_(0, 255)          # dirty whatever the allocator hands out next
        dec, n, ties = eng.decode_device(x, l)
        dec, n = dec.cpu(), n.cpu()
        bad = [i for i in range(B) if dec[i, :int(n[i])].tolist() != want[i]]
        if bad:
            nbad += 1
            for i in bad: seen[i] = seen.get(i, 0) + 1
    print(os.environ.get("E2E_CTC_LIB", "default")[-20:], "wide", os.environ.get("E2E_CTC_BEAM_WIDE"), (V, beam, T, B), "failing runs %d/%d" % (nbad, reps), seen, flush=True)
