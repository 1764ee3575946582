"""The reference's OWN unittest files, run unmodified against the drop-in ``pytorch_end2end`` package
(tests/test_ctc.py:22-191 -- five warp-ctc / TensorFlow known answers + gradcheck;
tests/test_ctc_decoder.py:44-166 -- greedy and prefix-beam known answers).

The two files are staged from /root/reference into oracle/_ref/tests/ by oracle/build_ref.py (git-ignored, shipped
to the GPU box); they are executed with the repository root first on PYTHONPATH, so
``from pytorch_end2end import CTCLoss, CTCDecoder`` resolves to the B200 engine.  The LM tests of the decoder
file need KenLM + the reference's ARPA fixture and stay out (north_star: KenLM decoding stays CPU code)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref", "tests")


def _run(module, tests=()):
    path = os.path.join(STAGED, module + ".py")
    if not os.path.exists(path):
        pytest.skip("reference test files not staged (oracle/build_ref.py stages them where /root/reference exists)")
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env.setdefault("OMP_NUM_THREADS", "4")
    names = [module + "." + t for t in tests] or [module]
    r = subprocess.run([sys.executable, "-m", "unittest", "-v"] + names, cwd=STAGED, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stderr


def test_reference_test_ctc_unmodified():
    out = _run("test_ctc")
    assert "Ran 6 tests" in out and "OK" in out, out[-2000:]


def test_reference_test_ctc_decoder_unmodified():
    """The greedy known answer (tests/test_ctc_decoder.py:44-59) and the three LM-free prefix-beam known answers
    (:86-166: greedy AND beam result of each case)."""
    out = _run("test_ctc_decoder", ["TestCTCDecoder.test_greedy_simple", "TestCTCDecoder.test_with_probs_sm",
                                    "TestCTCDecoder.test_with_probs_1", "TestCTCDecoder.test_with_probs_2"])
    assert "Ran 4 tests" in out and "OK" in out, out[-2000:]
