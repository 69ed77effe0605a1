"""The reference's own example, UNMODIFIED (examples/gpt-2/main-backend.cpp compiled with -DGGML_USE_CUDA into oracle/_ref/gpt-2-backend; its
ggml_backend_cuda_init(0) is this backend's alias), on a random-init GPT-2 117M Q4_0 model file in the reference's legacy ggml format
(oracle/make_gpt2_model.py): BASELINE.json configs[2] as north_star words it -- "the gpt-2-backend ... paths pick it up unchanged".  The same
source compiled for the reference CPU backend (oracle/_ref/gpt-2-backend-cpu) runs beside it.  Logit-level parity of this graph is
tests/test_gpu_gpt2_sched.py (arm b200); here the binary itself must load the file, tokenize, run prompt + decode on the GPU and exit 0."""
import re
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = ROOT / "oracle" / "_ref"
PROMPT = "the quick brown fox jumps over the lazy dog and keeps running through the forest until the night falls over the quiet hills of the north"


def run(exe, model, extra):
    p = subprocess.run([str(exe), "-m", str(model), "-p", PROMPT, "-n", "16", "-s", "7", "--top_k", "1", "-b", "128", "-t", "8"] + extra,
                       capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout, p.stderr


@pytest.mark.parametrize("qname", ["q4_0", "q8_0"])
def test_reference_gpt2_backend_example_runs_unmodified_on_b200(tmp_path, qname):
    gpu_exe, cpu_exe = REF / "gpt-2-backend", REF / "gpt-2-backend-cpu"
    assert gpu_exe.exists() and cpu_exe.exists(), "oracle/_ref/gpt-2-backend[-cpu] must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    model = tmp_path / f"gpt2-117m-{qname}.bin"
    subprocess.check_call([sys.executable, str(ROOT / "oracle" / "make_gpt2_model.py"), str(model), qname])
    rc, out, err = run(gpu_exe, model, ["-ngl", "1"])
    assert rc == 0, (out[-2000:], err[-2000:])
    assert "using CUDA backend" in err and "using CPU backend" not in err, err[-2000:]          # (the alias: ggml_backend_cuda_init -> B200)
    m = re.search(r"number of tokens in prompt = (\d+)", out)
    assert m and int(m.group(1)) >= 128, out[-2000:]
    gen = re.findall(r"<(\d+)>|\n", out.split(PROMPT)[-1].split("main:")[0])
    t = re.search(r"predict time =\s*([\d.]+) ms / ([\d.]+) ms per token", out)
    assert t, out[-2000:]
    rc2, out2, err2 = run(cpu_exe, model, [])
    assert rc2 == 0 and "using CPU backend" in err2
    t2 = re.search(r"predict time =\s*([\d.]+) ms / ([\d.]+) ms per token", out2)
    # greedy continuations: rounding-level logit differences may flip an argmax of a random-init model, so agreement is reported, not required
    g1 = out.split(PROMPT)[-1].split("\n\n")[0]
    g2 = out2.split(PROMPT)[-1].split("\n\n")[0]
    print(f"\n[gpt-2-backend {qname}] B200: {t.group(1)} ms total, {t.group(2)} ms/token; CPU (8 threads): {t2.group(1)} ms total, {t2.group(2)} ms/token; "
          f"same greedy continuation: {g1 == g2}")
    assert len(g1) > 0 and gen is not None


@pytest.mark.parametrize("ngl", [12, 6])
def test_reference_gpt2_sched_example_runs_unmodified_with_layers_on_b200(tmp_path, ngl):
    """examples/gpt-2/main-sched.cpp, unmodified: the reference's scheduler splits the graph between this backend (the last `ngl` layers'
    weights, and the KV cache / inputs when most layers are there) and the CPU backend by itself -- tensor copies between the two included"""
    exe = REF / "gpt-2-sched"
    assert exe.exists(), "oracle/_ref/gpt-2-sched must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    model = tmp_path / "gpt2-117m-q4_0.bin"
    subprocess.check_call([sys.executable, str(ROOT / "oracle" / "make_gpt2_model.py"), str(model), "q4_0"])
    rc, out, err = run(exe, model, ["-ngl", str(ngl)])
    assert rc == 0, (out[-2000:], err[-2000:])
    assert "using CUDA backend" in err, err[-2000:]
    t = re.search(r"predict time =\s*([\d.]+) ms / ([\d.]+) ms per token", out)
    assert t, out[-2000:]
    rc2, out2, _ = run(REF / "gpt-2-backend-cpu", model, [])
    g1 = out.split(PROMPT)[-1].split("\n\n")[0]
    g2 = out2.split(PROMPT)[-1].split("\n\n")[0]
    print(f"\n[gpt-2-sched -ngl {ngl}] {t.group(1)} ms total, {t.group(2)} ms/token; same greedy continuation as the CPU backend: {g1 == g2}")
    assert len(g1) > 0


def test_reference_gpt2_batched_example_runs_unmodified_on_b200(tmp_path):
    """examples/gpt-2/main-batched.cpp, unmodified, 4 parallel sequences: F16 KV cache (CPY F32 -> F16, MUL_MAT with an F16 src0), an explicit
    KQ mask broadcast over the heads (ADD), SOFT_MAX, a batched head"""
    exe, cpu_exe = REF / "gpt-2-batched", REF / "gpt-2-batched-cpu"
    assert exe.exists() and cpu_exe.exists(), "oracle/_ref/gpt-2-batched[-cpu] must be prebuilt (make -C oracle dropin) and travel with the snapshot"
    model = tmp_path / "gpt2-117m-q4_0.bin"
    subprocess.check_call([sys.executable, str(ROOT / "oracle" / "make_gpt2_model.py"), str(model), "q4_0"])
    outs = []
    for e, extra in ((exe, ["-ngl", "1"]), (cpu_exe, [])):
        p = subprocess.run([str(e), "-m", str(model), "-p", PROMPT, "-n", "16", "-s", "7", "--top_k", "1", "-b", "256", "-t", "8", "-np", "4"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
        outs.append(p.stdout + p.stderr)
    assert "using CUDA backend" in outs[0], outs[0][-2000:]
    t = [re.search(r"total time\s*=\s*([\d.]+) ms", o) for o in outs]
    seqs = [re.findall(r"sequence \d+:\n\n(.*)", o) for o in outs]
    print(f"\n[gpt-2-batched -np 4] B200 predict {t[0].group(1) if t[0] else '?'} ms, CPU predict {t[1].group(1) if t[1] else '?'} ms (prompt + 4 x 16 tokens); "
          f"sequences equal to the CPU backend's: {sum(a == b for a, b in zip(seqs[0], seqs[1]))} of {len(seqs[1])}")
    assert len(seqs[0]) == 4 and all(len(x) > 0 for x in seqs[0]), outs[0][-2000:]


def test_reference_simple_backend_example():
    """examples/simple/simple-backend.cpp, unmodified: its 4x2 by 2x3 F32 mul_mat on this backend prints the result its README gives"""
    exe = REF / "simple-backend"
    assert exe.exists(), "oracle/_ref/simple-backend must be prebuilt (make -C oracle dropin)"
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "using CUDA backend" in p.stderr, (p.stdout[-800:], p.stderr[-800:])
    nums = [float(v) for v in re.findall(r"-?\d+\.\d+", p.stdout.split("transposed result")[-1])]
    # (the example prints the 4 x 3 result with its own index arithmetic; the twelve values are what matters)
    assert sorted(nums) == sorted([60.0, 55.0, 50.0, 110.0, 90.0, 54.0, 54.0, 126.0, 42.0, 29.0, 28.0, 64.0]), p.stdout


def test_reference_test_mul_mat():
    """tests/test-mul-mat.cpp, unmodified: CONT + F32 MUL_MAT on this backend against the values the reference's test expects (exact)"""
    exe = REF / "test-mul-mat"
    assert exe.exists(), "oracle/_ref/test-mul-mat must be prebuilt (make -C oracle dropin)"
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    clean = re.sub(r"\x1b\[[0-9;]*m", "", p.stdout)
    assert p.returncode == 0, (clean[-800:], p.stderr[-800:])       # (its load_model(..., use_gpu = true) takes ggml_backend_cuda_init(0) = this backend)
    assert "ggml_mul_mat (64): PASSED" in clean and "FAILED" not in clean, clean[-1500:]
