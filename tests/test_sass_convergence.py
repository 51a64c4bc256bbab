"""The persistent loop of k_trace votes on "which lanes are idle" at its top, and stops when EVERY lane is (lcrt_kernels.cu).

nvcc 12.9 lowers that __ballot_sync(0xffffffff, ..) to a VOTE / VOTEU and, in some instantiations, puts no WARPSYNC between the
divergent primitive phase and the vote; the round-1 RayQuery instantiation, one of those, did not finish in half of its runs on
BASELINE C5's bounce rays (profiles/r2_trace_experiments.md §3).  The source therefore has two forms: instantiations that the
compiler gives a WARPSYNC.ALL on the back edge keep comparing against 0xffffffff (the headline closest-hit kernels: their code
must not change), all others compare against the lanes that actually take part in the vote.  Which instantiation gets what is the
compiler's decision, so this test reads the SASS of the library that ships and fails if a rebuild has moved a kernel from one
group to the other.  CPU only (cuobjdump)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "luisacompute_b200", "liblcb200_rt.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")


def _functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    return sorted({l.split("Function : ")[1].strip() for l in out.splitlines() if "Function : " in l and "k_trace" in l})


def _instructions(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def test_refill_vote_of_every_trace_kernel_is_safe():
    funs = _functions()
    assert len(funs) >= 8, funs        # closest (in place, FUSED_EACH, FUSED_BATCH, counted), any (2), RayQuery (2)
    report = {}
    for fun in funs:
        ins = _instructions(fun)
        top = next(a for a, t in ins if t.startswith("VOTE") and "!P0" in t)                    # ballot(!active)
        back = [a for a, t in ins if a > top and re.match(r"BRA 0x[0-9a-f]+$", t) and int(t.split("0x")[1], 16) <= top]
        assert back, fun
        index = {a: i for i, (a, _) in enumerate(ins)}
        synced = all(any(t.startswith("WARPSYNC.ALL") for _, t in ins[max(0, index[a] - 16):index[a]]) for a in back)
        i = index[top]
        # ballot(true) next to the refill ballot: "everybody" = whoever takes part in the vote
        takes_part = any(re.match(r"VOTEU?\.ANY U?R\d+, U?PT, U?PT$", t) for _, t in ins[max(0, i - 8):i + 8])
        report[fun] = (synced, takes_part)
        assert synced or takes_part, f"{fun}: no WARPSYNC.ALL on the back edge and the vote assumes a full warp"
    headline = [f for f in funs if "ILi0ELb0E" in f and ("NoCallbacksELi0E" in f or "NoCallbacksELi1E" in f)]
    assert len(headline) == 2
    for f in headline:
        # the kernels bench.py's `value` is measured on: converged by the compiler, and NOT paying for the extra vote
        assert report[f] == (True, False), (f, report[f])
