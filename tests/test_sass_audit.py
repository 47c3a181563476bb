"""CPU (needs cuobjdump): the exact kernels must not contain contracted multiply-adds.

ptxas contracts packed mul+add into FFMA2 unless the multiply is written fma(a, b, +0) (DESIGN.md
section 2), so every FFMA2 in the front-end / stage FIR kernels must have RZ as its addend, and the
scalar FFMA instructions that remain must belong to the IEEE reciprocal/divide expansions."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "singlecarrier_b200", "libsinglecarrier_b200.so")


def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", LIB], check=True, capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            funcs[cur].append(ln.split("*/", 1)[1].split("/*")[0].strip())
    return funcs


def test_no_contracted_fma_in_exact_kernels():
    funcs = sass()
    assert any("sm_100a" in x or True for x in funcs)
    seen = 0
    for name, ins in funcs.items():
        if not any(k in name for k in ("frontend_kernel", "fir_batch10_kernel", "packet_fir_kernel")):
            continue
        if re.search(r"fir_batch(10)?_kernelILb[01]ELb1E", name):
            # the explicitly named tolerance mode (SC_FIR_FAST) is the one place where contraction is wanted
            assert any("FFMA2" in i and not i.rstrip(" ;").endswith(("RZ", "RZ.F32")) for i in ins), name
            continue
        seen += 1
        packed = [i for i in ins if "FFMA2" in i]
        assert len(packed) >= (98 if "packet_fir" in name else 200), name
        for i in packed:
            assert re.search(r",\s*RZ(\.F32)?\s*;?$", i.rstrip(" ;") + ";") or i.rstrip(" ;").endswith("RZ.F32") \
                or i.rstrip(" ;").endswith("RZ"), (name, i)
        assert sum("FADD2" in i for i in ins) >= len(packed)
        scalar_ffma = [i for i in ins if re.match(r"(@!?P\d+\s+)?FFMA\b", i)]
        if re.search(r"frontend_kernelILb[01]ELb0ELb1E", name):
            # tensor-core proposed search: HMMA tiles, and the only scalar FFMAs are the correctly rounded sqrtf() of the
            # candidate bound (an inequality, not part of the reference's arithmetic)
            assert sum("HMMA" in i for i in ins) == 72, name
            assert len(scalar_ffma) <= 4, (name, scalar_ffma)
        else:
            assert not scalar_ffma, name                                              # no scalar FFMA at all here
            assert not [i for i in ins if "HMMA" in i], name
    assert seen >= 9
    for name, ins in funcs.items():
        if "search_batch_kernel" in name:                      # pure adds: no multiply-add of any kind
            assert not [i for i in ins if "FFMA" in i], name
            assert sum(bool(re.match(r"(@!?P\d+\s+)?FADD\b", i)) for i in ins) >= 1024
    for name, ins in funcs.items():
        if "track_kernel" in name or "track_window_kernel" in name:
            # 5 reciprocals per step x 2 loops (+ slow paths); anything beyond that would be a contraction
            n_ffma = sum(bool(re.match(r"(@!?P\d+\s+)?FFMA\b", i)) for i in ins)
            n_rcp = sum("MUFU.RCP" in i for i in ins)
            assert n_rcp >= 10 and n_ffma <= 4 * n_rcp, (name, n_ffma, n_rcp)
            assert sum(bool(re.match(r"(@!?P\d+\s+)?FMUL\b", i)) for i in ins) > 300
            assert not [i for i in ins if "FFMA2" in i]


def test_compiled_for_sm_100a_only():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-lelf", LIB], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_tcgen05_search_kernel_uses_the_blackwell_units():
    """search_umma_batch_kernel (sc_search_umma.cu): the proposer's GEMM is tcgen05.mma with A and D in tensor memory
    (UTCHMMA, LDTM / STTM, the allocator's UTCATOMSWS), its completion a tcgen05.commit on an mbarrier (UTCBAR), the
    windows arrive by TMA bulk copies (UBLKCP) -- and the verifier's sums contain no multiply-add."""
    funcs = sass()
    names = [n for n in funcs if "search_umma_batch_kernel" in n]
    assert len(names) == 1, names
    ins = funcs[names[0]]
    # UBLKCP twice: the windows (TMA warp) and the verifier's gather of a candidate's symbols
    for mnem, least in (("UTCHMMA", 1), ("LDTM", 4), ("STTM", 1), ("UTCBAR", 1), ("UBLKCP", 2), ("UTCATOMSWS", 2),
                        ("SYNCS", 8), ("REDUX", 16)):
        assert sum(mnem in i for i in ins) >= least, (mnem, sum(mnem in i for i in ins))
    flush = [n for n in funcs if "su_flush_queue" in n]     # the queued second candidates: cp.async gather, out of line
    assert flush and sum("LDGSTS" in i for i in funcs[flush[0]]) >= 33
    assert not [i for i in ins if "FFMA" in i or "HMMA.16816" in i]
    assert sum(bool(re.match(r"(@!?P\d+\s+)?FADD\b", i)) for i in ins) >= 256


def test_fused_tcgen05_frontend_uses_the_blackwell_units_and_no_contraction():
    """frontend_umma_kernel (sc_frontend_umma.cu, SC_FE_SEARCH_TCGEN05): the FIR is the front-end's (packed multiplies
    with a zero addend, separate packed adds -- nothing contracted); the proposer is tcgen05.mma with both operands
    through shared-memory descriptors and D in tensor memory (UTCHMMA, LDTM, the allocator's UTCATOMSWS), completion by
    tcgen05.commit (UTCBAR); the Toeplitz master and the call's phasor table arrive by TMA bulk copies (UBLKCP); the
    verifier's sums are plain adds."""
    funcs = sass()
    names = [n for n in funcs if "frontend_umma_kernel" in n]
    assert len(names) == 2, names                               # narrow / wide tap table
    for n in names:
        ins = funcs[n]
        for mnem, least in (("UTCHMMA", 1), ("LDTM", 4), ("UTCBAR", 1), ("UBLKCP", 2), ("UTCATOMSWS", 2), ("SYNCS", 6),
                            ("REDUX", 16)):
            assert sum(mnem in i for i in ins) >= least, (n, mnem, sum(mnem in i for i in ins))
        packed = [i for i in ins if "FFMA2" in i]
        assert len(packed) >= 400, n
        for i in packed:
            assert i.rstrip(" ;").endswith(("RZ", "RZ.F32")), (n, i)
        assert sum("FADD2" in i for i in ins) >= len(packed)
        # the only scalar multiply-adds allowed are those of the bound's sqrt / none at all (sqrt.approx is a MUFU)
        assert not [i for i in ins if re.match(r"(@!?P\d+\s+)?FFMA\b", i)], n
        assert not [i for i in ins if "HMMA.16816" in i], n
        assert sum(bool(re.match(r"(@!?P\d+\s+)?FADD\b", i)) for i in ins) >= 256
