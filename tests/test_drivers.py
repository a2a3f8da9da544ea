"""The five C drivers (host/*.c -> bin/<format>): exit-code contract without a GPU (CPU test) and,
on the GPU, the reference's stdout contract line by line next to the unmodified reference binary
(oracle/_ref/bin/<format>, fake OpenCL) run in the same directory on the same files."""
import os
import re
import subprocess
from pathlib import Path

import pytest

from conftest import ROOT
from oracle import binding as O

HOST = ROOT / "opencl-spmv-algorithms_b200" / "host"
FORMATS = ("coo", "csr", "ell", "sigma_c", "cmrs")


def build_drivers():
    subprocess.run(["make", "-C", str(HOST), "all"], check=True, capture_output=True)
    return HOST / "bin"


def run(binary, cwd, *args):
    return subprocess.run([str(binary), *args], cwd=cwd, capture_output=True, text=True, timeout=600)


def test_drivers_build_with_reference_warning_flags():
    bins = build_drivers()
    for f in FORMATS:
        assert (bins / f).exists()
    flags = (HOST / "Makefile").read_text()
    assert "-Wall -Werror -Wshadow" in flags and "-fopenmp" in flags   # reference Makefile:18


def test_no_gpu_exit_code(tmp_path):
    """Without a device every driver stops with OpenCLDeviceError (1), like the reference when
    get_device_ids fails (csr.c:25-28) -- never a CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    bins = build_drivers()
    for f in FORMATS:
        p = run(bins / f, tmp_path)
        assert p.returncode == 1, (f, p.stdout, p.stderr)
        assert "CPU calculations" not in p.stdout
    assert run(bins / "csr", tmp_path, "--bogus").returncode == 4      # OtherError
    # the iterated mode: argument errors are OtherError, a missing device is OpenCLDeviceError -- and the
    # synthetic path must not need a file
    assert run(bins / "csr", tmp_path, "--gpus", "2").returncode == 4             # --gpus without --iters
    assert run(bins / "sigma_c", tmp_path, "--iters", "5", "--sync", "smoke").returncode == 4
    assert run(bins / "ell", tmp_path, "--iters", "5").returncode == 4            # csr and sigma_c only
    p = run(bins / "sigma_c", tmp_path, "--synthetic", "laplace7:8x8x8", "--iters", "5", "--gpus", "2", "--json")
    assert p.returncode == 1 and "No CUDA devices found" in p.stdout, (p.stdout, p.stderr)


def test_fast_parallel_parse_equals_fscanf_parse(cant_dir, tmp_path):
    """The drivers' one-pass, multi-threaded text parse (host/src/driver_common.c: read_entries)
    must give bit-identical triples to the reference-style fscanf("%d %d %lg") parse -- on a large
    well-formed file (parallel path) and on files whose entries are NOT one per line (it must notice
    and fall back to the line-agnostic serial parse)."""
    import numpy as np
    bins = build_drivers()

    def parse(path, threads):
        out = tmp_path / "p.bin"
        p = subprocess.run([str(bins / "mtx_parse"), str(path), str(out)], capture_output=True, text=True,
                           env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
        assert p.returncode == 0, p.stderr
        n_rows, n_cols, nnz = (int(t) for t in p.stdout.split()[:3])
        raw = out.read_bytes()
        rows = np.frombuffer(raw, np.int32, nnz, 0)
        cols = np.frombuffer(raw, np.int32, nnz, 4 * nnz)
        vals = np.frombuffer(raw, np.float64, nnz, 8 * nnz)
        return n_rows, n_cols, rows, cols, vals

    big = cant_dir / "databases" / "cant.mtx"
    ref = O.read_mtx(big)
    for threads in (1, 8):
        got = parse(big, threads)
        assert got[:2] == ref[:2]
        for a, b in zip(got[2:], ref[2:]):
            assert a.tobytes() == b.tobytes()
    # entries wrapped over lines / several per line: legal for fscanf, so legal here
    n = 40000
    rng = np.random.default_rng(0)
    r, c, v = rng.integers(1, 500, n), rng.integers(1, 500, n), np.round(rng.uniform(-9, 9, n), 7)
    weird = tmp_path / "weird.mtx"
    with open(weird, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% odd but scanf-legal layout\n\n" + f"500 500 {n}\n")
        for i in range(0, n, 2):
            f.write(f"{r[i]} {c[i]}\n{float(v[i])!r} {r[i + 1]} {c[i + 1]} {float(v[i + 1])!r}\n")
    ref = O.read_mtx(weird)
    got = parse(weird, 8)
    for a, b in zip(got[2:], ref[2:]):
        assert a.tobytes() == b.tobytes()


def test_number_scanners_equal_fscanf_on_awkward_spellings(tmp_path):
    """The drivers' own integer / double scanners (fast_int, fast_double: Clinger's fast path, strtod for the
    rest) against the reference-style fscanf("%d %d %lg") on every spelling %lg accepts: 1..25 significant
    digits, decimal exponents inside and outside [-22, 22], leading zeros, bare '.5' / '5.', signs, signed
    zero, overflow to inf, underflow to zero and to denormals, inf / nan, hex floats; row / column numbers with
    a '+' or leading zeros.  Bit-identical triples, serial and parallel parse, mapped and copied file."""
    import numpy as np
    bins = build_drivers()
    rng = np.random.default_rng(12)
    special = ["0", "-0", "0.0", "-0.0e7", ".5", "5.", "+7", "-.25e1", "1e5", "1E-5", "1.5e+300", "4.9e-324", "2.2250738585072014e-308",
               "1e-400", "-1e400", "inf", "-inf", "nan", "0x1.8p3", "-0X10", "123456789012345678901234", "0.000000000000000000000123",
               "9007199254740993", "9007199254740992e1", "1e22", "1e23", "9007199254740991e22", "8.5e-22", "8.5e-23",
               "179769313486231570814527423731704356798070567525844996598917476803157260780028538760589558632766878171540458953"
               "514382464234321326889464182768467546703537516986049910576551282076245490090389328944075868508455133942304583236"
               "90322294816580855933212334827479782620414472316873817718091929988125040402618412485836800000000000000000000000"]
    vals = list(special)
    while len(vals) < 20000:
        digits = int(rng.integers(1, 26))
        mant = "".join(str(d) for d in rng.integers(0, 10, digits))
        cut = int(rng.integers(0, digits + 1))
        text = (mant[:cut] or ("" if rng.random() < 0.3 else "0")) + "." + mant[cut:] if rng.random() < 0.8 else mant
        if text == ".":
            text = "0."
        if rng.random() < 0.5:
            text += ("e", "E")[int(rng.integers(0, 2))] + ("", "+", "-")[int(rng.integers(0, 3))] + str(int(rng.integers(0, 40)))
        vals.append(("-", "+", "")[int(rng.integers(0, 3))] + text)
    n = len(vals)
    r, c = rng.integers(1, 900, n), rng.integers(1, 900, n)

    def write(path, pad_to_page):
        body = "%%MatrixMarket matrix coordinate real general\n" + f"900 900 {n}\n"
        for i in range(n):
            rr = f"+{r[i]}" if i % 7 == 0 else (f"00{r[i]}" if i % 11 == 0 else str(r[i]))
            body += f"{rr} {c[i]} {vals[i]}\n"
        if pad_to_page:
            body += " " * ((-len(body) - 1) % 4096) + "\n"
            assert len(body) % 4096 == 0
        else:
            assert len(body) % 4096 != 0
        path.write_text(body)

    for pad in (False, True):
        path = tmp_path / f"awkward{int(pad)}.mtx"
        write(path, pad)
        ref = O.read_mtx(path)
        assert np.isnan(ref[4]).sum() >= 1 and np.isinf(ref[4]).sum() >= 3      # the reference-style reader took them all
        for threads in (1, 8):
            out = tmp_path / "p.bin"
            p = subprocess.run([str(bins / "mtx_parse"), str(path), str(out)], capture_output=True, text=True,
                               env=dict(os.environ, OMP_NUM_THREADS=str(threads), B200_PARSE_TIMING="1"))
            assert p.returncode == 0, p.stderr
            assert ("mapped" if not pad else "read") in p.stderr, p.stderr
            raw = out.read_bytes()
            got = (np.frombuffer(raw, np.int32, n, 0), np.frombuffer(raw, np.int32, n, 4 * n), np.frombuffer(raw, np.float64, n, 8 * n))
            for a, b in zip(got, ref[2:]):
                assert a.tobytes() == b.tobytes(), (pad, threads)


def test_expand_symmetric_equals_the_full_matrix(tmp_path):
    """--expand-symmetric (new, optional): the lower triangle of a file whose banner says `symmetric`,
    mirrored and sorted by (row, column), must equal the row-sorted FULL matrix bit for bit (the
    generator's values are symmetric by construction); skew-symmetric negates the mirrored entries;
    a `general` banner is left untouched (the reference's reading, and the default without the flag)."""
    import numpy as np
    from conftest import gen_mtx_tool, write_mtx
    bins = build_drivers()
    g = ["--grid", "6", "5", "9", "--dof", "3", "--order", "row"]
    lower, full = tmp_path / "lower.mtx", tmp_path / "full.mtx"
    subprocess.run([str(gen_mtx_tool()), *g, "--tri", "lower", "--banner", "symmetric", "--out", str(lower)], check=True)
    subprocess.run([str(gen_mtx_tool()), *g, "--tri", "full", "--out", str(full)], check=True)

    def parse(path, *flags):
        out = tmp_path / "p.bin"
        p = subprocess.run([str(bins / "mtx_parse"), str(path), str(out), *flags], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        n_rows, n_cols, nnz = (int(t) for t in p.stdout.split()[:3])
        raw = out.read_bytes()
        return (n_rows, n_cols, np.frombuffer(raw, np.int32, nnz, 0), np.frombuffer(raw, np.int32, nnz, 4 * nnz),
                np.frombuffer(raw, np.float64, nnz, 8 * nnz))

    want = O.read_mtx(full)
    got = parse(lower, "--expand-symmetric")
    assert got[:2] == want[:2] and got[2].size == want[2].size > O.read_mtx(lower)[2].size
    for a, b in zip(got[2:], want[2:]):
        assert a.tobytes() == b.tobytes()
    # without the flag: the lower triangle as stored (what the reference multiplies by)
    asis = parse(lower)
    for a, b in zip(asis[2:], O.read_mtx(lower)[2:]):
        assert a.tobytes() == b.tobytes()
    # general banner: the flag changes nothing
    same = parse(full, "--expand-symmetric")
    for a, b in zip(same[2:], want[2:]):
        assert a.tobytes() == b.tobytes()
    # skew-symmetric: mirrored entries negated, result sorted by (row, col)
    rows = np.array([1, 2, 2, 3, 3], np.int32)
    cols = np.array([0, 0, 1, 1, 2], np.int32)
    vals = np.array([1.5, -2.0, 3.25, 4.0, -5.5])
    skew = tmp_path / "skew.mtx"
    write_mtx(skew, 4, 4, rows, cols, vals, banner="skew-symmetric")
    _, _, r, c, v = parse(skew, "--expand-symmetric")
    dense = np.zeros((4, 4))
    dense[rows, cols] = vals
    dense[cols, rows] = -vals
    rr, cc = np.nonzero(dense)
    assert np.array_equal(r, rr) and np.array_equal(c, cc) and np.array_equal(v, dense[rr, cc])


def test_binary_cache_of_parsed_triples(tmp_path):
    """--cache (new, optional): the first load writes <file>.b200cache, later loads of an unchanged file
    read it and give the same bytes; a changed source (size or mtime) invalidates it; the cached
    triples are the ones BEFORE symmetric expansion, so both options compose."""
    import numpy as np
    from conftest import gen_mtx_tool
    bins = build_drivers()
    src = tmp_path / "m.mtx"
    g = ["--grid", "6", "5", "9", "--dof", "3", "--order", "row", "--tri", "lower", "--banner", "symmetric"]
    subprocess.run([str(gen_mtx_tool()), *g, "--out", str(src)], check=True)

    def parse(*flags):
        out = tmp_path / "p.bin"
        p = subprocess.run([str(bins / "mtx_parse"), str(src), str(out), *flags], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        return p.stdout.split()[:3], out.read_bytes()

    plain = parse()
    cache = tmp_path / "m.mtx.b200cache"
    assert not cache.exists()                                  # nothing is written without the flag
    first = parse("--cache")
    assert cache.exists() and first == plain
    stamp = cache.stat().st_mtime_ns
    again = parse("--cache")
    assert again == plain and cache.stat().st_mtime_ns == stamp   # served from the cache, not rewritten
    assert parse("--cache", "--expand-symmetric") == parse("--expand-symmetric")
    # a corrupted cache body with a valid header would go unnoticed -- the header is the contract -- but a
    # changed SOURCE must invalidate: append a comment-free duplicate entry and fix the size line
    lines = src.read_text().splitlines()
    n_rows, n_cols, nnz = (int(t) for t in lines[2].split())
    lines[2] = f"{n_rows} {n_cols} {nnz + 1}"
    lines.append(lines[-1])
    src.write_text("\n".join(lines) + "\n")
    changed = parse("--cache")
    assert int(changed[0][2]) == nnz + 1 and changed == parse()
    assert parse("--cache") == changed
    # garbage in place of the cache is ignored and replaced
    cache.write_bytes(b"not a cache")
    assert parse("--cache") == changed and cache.stat().st_size > 64


TIMING = re.compile(r"^(Your calculations took|Number of operations \d+, PERFORMANCE|GBytes transferred)")


def shape(stdout):
    """stdout with the timing numbers blanked: what must match the reference byte for byte."""
    out = []
    for line in stdout.splitlines():
        m = re.match(r"^(Number of operations \d+), PERFORMANCE", line)
        if m:
            out.append(m.group(1))
        elif line.startswith("Your calculations took"):
            out.append("Your calculations took")
        elif line.startswith("GBytes transferred to processor"):
            out.append(" ".join(line.split()[:7]))   # the two byte counts are deterministic
        else:
            out.append(line)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FORMATS)
def test_driver_stdout_matches_reference(fem_small_dir, fmt):
    bins = build_drivers()
    ours = run(bins / fmt, fem_small_dir)
    assert ours.returncode == 0, ours.stdout + ours.stderr
    assert "result is ok" in ours.stdout and "result is wrong" not in ours.stdout
    if fmt != "sigma_c":
        assert "cpu result is ok" in ours.stdout
    if not O.ref_available():
        pytest.skip("oracle/_ref not built")
    O.prepare_ref_workdir(fem_small_dir)
    ref = run(O.REF_DIR / "bin" / fmt, fem_small_dir)
    assert ref.returncode == 0
    # the fake OpenCL runtime computes nothing, so the reference's GPU section says "wrong"
    ref_lines = [l for l in shape(ref.stdout) if not l.startswith("wrong value at index")]
    ref_lines = ["result is ok" if l == "result is wrong" else l for l in ref_lines]
    assert shape(ours.stdout) == ref_lines


@pytest.mark.gpu
def test_driver_options_and_errors(fem_small_dir, tmp_path):
    bins = build_drivers()
    assert run(bins / "csr", tmp_path).returncode == 3                  # FileError: no databases/
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n")
    assert run(bins / "csr", tmp_path, "--matrix", str(bad)).returncode == 3
    mtx = str(fem_small_dir / "databases" / "cant-sorted.mtx")
    for fmt in ("csr", "ell", "sigma_c", "cmrs"):
        p = run(bins / fmt, tmp_path, "--matrix", mtx, "--dtype", "f32", "--reps", "3")
        assert p.returncode == 0 and "result is ok" in p.stdout, (fmt, p.stdout, p.stderr)
    p = run(bins / "coo", tmp_path, "--matrix", str(fem_small_dir / "databases" / "cant.mtx"), "--dtype", "f32")
    assert p.returncode == 0 and "result is ok" in p.stdout
    for sigma in ("32", "128", "1000"):
        p = run(bins / "sigma_c", tmp_path, "--matrix", mtx, "--sigma", sigma)
        assert p.returncode == 0 and "result is ok" in p.stdout, (sigma, p.stdout)
    for layout in ("--rowmajor", "--colmajor"):
        p = run(bins / "ell", tmp_path, "--matrix", mtx, layout)
        assert p.returncode == 0 and "result is ok" in p.stdout, (layout, p.stdout)


SHIM_BINS = ROOT / "oracle" / "_ref" / "bin_b200"


def test_shim_exports_the_opencl_symbols_the_reference_links():
    """libOpenCL_b200.so must export the 21 OpenCL entry points the reference drivers call."""
    lib = ROOT / "opencl-spmv-algorithms_b200" / "lib" / "libOpenCL_b200.so"
    subprocess.run(["make", "-C", str(ROOT / "opencl-spmv-algorithms_b200" / "shim"), "all"], check=True,
                   capture_output=True)
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (cl[A-Za-z]+)", out))
    needed = {"clGetPlatformIDs", "clGetDeviceIDs", "clCreateContext", "clCreateCommandQueueWithProperties",
              "clCreateBuffer", "clCreateProgramWithSource", "clBuildProgram", "clGetProgramBuildInfo",
              "clCreateKernel", "clSetKernelArg", "clEnqueueWriteBuffer", "clEnqueueReadBuffer",
              "clEnqueueNDRangeKernel", "clWaitForEvents", "clFinish", "clFlush", "clReleaseMemObject",
              "clReleaseCommandQueue", "clReleaseKernel", "clReleaseProgram", "clReleaseContext"}
    assert needed <= exported, needed - exported


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", FORMATS)
def test_unmodified_reference_binary_runs_on_b200(fem_small_dir, fmt):
    """the reference's own, unmodified driver (built from /root/reference against
    shim/CL/cl.h, linked with libOpenCL_b200.so instead of libOpenCL.so) runs its SpMV on the B200
    kernels and its OWN check_result accepts the GPU result: 'result is ok', exit code 0."""
    if not (SHIM_BINS / fmt).exists():
        pytest.skip("oracle/_ref/bin_b200 not built (needs the reference sources at build time)")
    O.prepare_ref_workdir(fem_small_dir)          # kernels/*.cl placeholders: the driver insists on reading them
    p = run(SHIM_BINS / fmt, fem_small_dir)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert "result is ok" in lines and "result is wrong" not in lines, p.stdout
    if fmt != "sigma_c":
        assert "cpu result is ok" in lines


@pytest.mark.gpu
def test_unmodified_reference_csr_on_full_cant_shape(cant_dir):
    if not (SHIM_BINS / "csr").exists():
        pytest.skip("oracle/_ref/bin_b200 not built")
    O.prepare_ref_workdir(cant_dir)
    p = run(SHIM_BINS / "csr", cant_dir)
    assert p.returncode == 0 and "result is ok" in p.stdout.splitlines(), p.stdout + p.stderr


def _gpu_count():
    import ctypes
    try:
        from __graft_entry__ import load_package
        n = ctypes.c_int(0)
        return n.value if load_package().lib().b200_get_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [1, 2, 4])
def test_driver_iterated_mode(fem_small_dir, tmp_path, gpus):
    """--iters K [--gpus N] [--json]: the power iteration issued from C (host thread per device, the
    library's NCCL communicator and iterator).  sigma_c = fused SELL kernel + halo exchange, csr =
    SpMV + ncclAllGather; both must agree with the driver's own serial CPU iteration (`ok`), with each
    other, and across device counts; --synthetic generates the Laplacian on the devices."""
    import json
    if _gpu_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    bins = build_drivers()
    mtx = str(fem_small_dir / "databases" / "cant-sorted.mtx")
    norms = {}
    for fmt in ("sigma_c", "csr"):
        p = run(bins / fmt, tmp_path, "--matrix", mtx, "--iters", "40", "--gpus", str(gpus), "--json")
        assert p.returncode == 0, p.stdout + p.stderr
        rec = json.loads(p.stdout.strip().splitlines()[-1])
        assert rec["ok"] is True and rec["checked"] is True and rec["gpus"] == gpus and rec["iters"] == 40
        assert abs(rec["norm"] - rec["cpu_norm"]) <= 1e-10 * abs(rec["cpu_norm"])
        norms[fmt] = rec["norm"]
    assert abs(norms["sigma_c"] - norms["csr"]) <= 1e-10 * abs(norms["csr"])
    # text output keeps the reference's vocabulary
    p = run(bins / "sigma_c", tmp_path, "--matrix", mtx, "--iters", "12", "--gpus", str(gpus))
    assert p.returncode == 0 and "result is ok" in p.stdout and "PERFORMANCE" in p.stdout, p.stdout + p.stderr
    # the synthetic Laplacian: same eigenvalue estimate whatever the device count (power iteration from
    # the same seeded x0), and equal to the one-device run of the other driver
    p = run(bins / "sigma_c", tmp_path, "--synthetic", "laplace7:20x16x27", "--iters", "30", "--gpus", str(gpus), "--json")
    assert p.returncode == 0, p.stdout + p.stderr
    a = json.loads(p.stdout.strip().splitlines()[-1])
    p = run(bins / "csr", tmp_path, "--synthetic", "laplace7:20x16x27", "--iters", "30", "--json")
    assert p.returncode == 0, p.stdout + p.stderr
    b = json.loads(p.stdout.strip().splitlines()[-1])
    assert a["rows"] == b["rows"] == 20 * 16 * 27 and a["nnz"] == b["nnz"]
    assert abs(a["norm"] - b["norm"]) <= 1e-10 * b["norm"] and 0 < a["norm"] < 12.0
    # --sync mcast: the all-reduce + barrier through NVSwitch multicast; same eigenvalue estimate, or a clean
    # refusal (exit code 1) where the box cannot set a multicast object up
    if gpus > 1:
        p = run(bins / "sigma_c", tmp_path, "--synthetic", "laplace7:20x16x27", "--iters", "30", "--gpus", str(gpus),
                "--sync", "mcast", "--json")
        if p.returncode == 0:
            c = json.loads(p.stdout.strip().splitlines()[-1])
            assert "multicast" in c["mode"] and abs(c["norm"] - b["norm"]) <= 1e-10 * b["norm"], c
        else:
            assert p.returncode == 1 and "multicast is not available" in p.stderr, p.stdout + p.stderr
    # argument errors
    assert run(bins / "coo", tmp_path, "--iters", "3").returncode == 4
    assert run(bins / "csr", tmp_path, "--gpus", "2").returncode == 4          # --gpus without --iters
    assert run(bins / "csr", tmp_path, "--iters", "3", "--gpus", "64").returncode == 4
