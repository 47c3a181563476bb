"""CPU: the restatement against the REFERENCE's own object code on fresh random inputs.  Only runs where
oracle/_ref/libsc_ref.so exists (built from /root/reference by oracle/Makefile; it travels to the GPU
box as a prebuilt file)."""
import numpy as np

from helpers import synth_streams


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_rx_random_streams(oracle, ref):
    rng = np.random.default_rng(99)
    samples = synth_streams(oracle, rng, 24, 10)
    noise = [rng.integers(-a, a + 1, 1880 * 8).astype(np.int16) for a in (3, 100, 3000, 32767)]
    for x in list(samples) + noise:
        b1, s1 = oracle.run_stream(x)
        b2, s2 = ref.run_stream(x)
        assert np.array_equal(s1["valid"], s2["valid"]) and np.array_equal(s1["max_index"], s2["max_index"])
        assert np.array_equal(s1["rx_timing"], s2["rx_timing"])
        assert same(s1["max_value"], s2["max_value"]) and same(s1["eq_coeff"], s2["eq_coeff"])
        v = s2["valid"].astype(bool)
        assert np.array_equal(b1[v], b2[v]) and np.array_equal(s1["matches"][v], s2["matches"][v])
        assert same(s1["cost"][v], s2["mean"][v])


def test_wide_filter(oracle, ref):
    rng = np.random.default_rng(5)
    x = synth_streams(oracle, rng, 3, 8)
    for s in range(3):
        b1, s1 = oracle.run_stream(x[s], wide=True)
        b2, s2 = ref.run_stream(x[s], wide=True)
        assert np.array_equal(s1["valid"], s2["valid"]) and same(s1["eq_coeff"], s2["eq_coeff"])


def test_stage_taps_filtered_and_decimated(oracle, ref):
    """Filtered samples and decimated symbols after a call: bit-identical."""
    rng = np.random.default_rng(6)
    x = synth_streams(oracle, rng, 1, 6)[0]
    ref.reset()
    st = oracle.new_state()
    for n in range(6):
        fr = x[n * 1880:(n + 1) * 1880]
        ref.rx_frame(fr)
        oracle.rx_frame(st, fr)
        mine = np.frombuffer(st, np.float32, count=2 * 3760).view(np.complex64)
        dec = np.frombuffer(st, np.float32, count=2 * 752, offset=3760 * 8).view(np.complex64)
        assert same(mine[:1880].view(np.float32), ref.filtered().view(np.float32))
        assert same(dec.view(np.float32), ref.decimated().view(np.float32))


def test_tx_random(oracle, ref):
    rng = np.random.default_rng(7)
    st = oracle.new_state()
    ref.reset()
    for _ in range(5):
        assert np.array_equal(oracle.tx_preamble(st), ref.tx_preamble())
        for _ in range(8):
            b = rng.integers(0, 2, 62).astype(np.uint8)
            assert np.array_equal(oracle.tx_data(st, b), ref.tx_data(b))


def test_f3_fix_removes_the_undefined_behaviour(tmp_path, gold):
    """the reference's decimated_frame[562] is written up to index 751.  Built under
    ASan/UBSan, the reference with the one-line fix the oracle build applies is clean on the shipped file,
    and without it the sanitizer reports the out-of-bounds access (so the fix is necessary and sufficient)."""
    import os
    import shutil
    import subprocess
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "qpsk.c")) or shutil.which("gcc") is None:
        import pytest
        pytest.skip("needs /root/reference")
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    pre, post = open(os.path.join(here, "ref_harness_pre.h")).read(), open(os.path.join(here, "ref_harness_post.c")).read()
    src = open(os.path.join(ref, "src", "qpsk.c")).read()
    drv = tmp_path / "drv.c"
    drv.write_text('#include <stdio.h>\n#include <stdint.h>\n'
                   'void ref_run_stream(const int16_t in[], int n_frames, int wide, uint8_t bits[], void *stats);\n'
                   'int main(int c, char **v){ FILE *f = fopen(v[1], "rb"); static int16_t x[30000]; size_t n = fread(x, 2, 30000, f);'
                   ' static uint8_t bits[20 * 62]; ref_run_stream(x, (int) (n / 1880), 0, bits, NULL); puts("done"); return 0; }\n')
    raw = tmp_path / "in.raw"
    gold("preamble_qpsk_8k.raw").tofile(raw)
    others = [os.path.join(ref, "src", f + ".c") for f in ("fir", "kalman", "equalizer", "scramble", "constants", "fft")]
    out = {}
    for name, body in (("fixed", src.replace("decimated_frame[562]", "decimated_frame[752]")), ("unfixed", src)):
        c = tmp_path / (name + ".c")
        c.write_text(pre + body + post)
        exe = tmp_path / name
        r = subprocess.run(["gcc", "-std=gnu11", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-w",
                            "-I", os.path.join(ref, "headers"), str(c), str(drv)] + others + ["-lm", "-o", str(exe)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            import pytest
            pytest.skip("sanitizer runtime not available: " + r.stderr[-200:])
        out[name] = subprocess.run([str(exe), str(raw)], capture_output=True, text=True, timeout=120)
    assert "done" in out["fixed"].stdout and "runtime error" not in out["fixed"].stderr and "ERROR" not in out["fixed"].stderr
    assert "out of bounds" in out["unfixed"].stderr


def test_oracle_restatement_is_sanitizer_clean(tmp_path, oracle):
    """The restatement itself under ASan/UBSan on a noisy loop-back stream."""
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    drv = tmp_path / "drv.c"
    drv.write_text('#include <stdio.h>\n#include <stdint.h>\n#include "sc_oracle.h"\n'
                   'int main(int c, char **v){ FILE *f = fopen(v[1], "rb"); static int16_t x[40000]; size_t n = fread(x, 2, 40000, f);'
                   ' static uint8_t bits[32 * 62]; static sco_frame_stats st[32]; sco_run_stream(x, (int) (n / 1880), 0, 0.0f, bits, st);'
                   ' int v2 = 0; for (int i = 0; i < (int) (n / 1880); i++) v2 += st[i].valid; printf("done %d\\n", v2); return 0; }\n')
    rng = np.random.default_rng(3)
    x = synth_streams(oracle, rng, 1, 20, noise_levels=(500.0,))[0]
    raw = tmp_path / "in.raw"
    x.tofile(raw)
    exe = tmp_path / "o"
    r = subprocess.run(["gcc", "-std=gnu11", "-O1", "-g", "-ffp-contract=off", "-fsanitize=address,undefined", "-I", here,
                        os.path.join(here, "sc_oracle.c"), str(drv), "-lm", "-o", str(exe)], capture_output=True, text=True)
    if r.returncode != 0:
        import pytest
        pytest.skip("sanitizer runtime not available")
    run = subprocess.run([str(exe), str(raw)], capture_output=True, text=True, timeout=120)
    assert run.stdout.startswith("done") and "runtime error" not in run.stderr and "ERROR" not in run.stderr
