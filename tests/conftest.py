"""Shared fixtures.  `-m "not gpu"` runs everywhere (oracle vs the committed reference vectors, host
logic, ABI export checks, the device algorithm under the CPU emulation of tests/emul); `-m gpu`
tests are the parity tests proper and call the CUDA library through its C-ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from genomeassembler_dev_b200 import tables  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with gpurun)")


@pytest.fixture(scope="session")
def kmers():
    return tables.all_kmer_strings()


@pytest.fixture(scope="session")
def raw_prob():
    return tables.load_raw()


@pytest.fixture(scope="session")
def prob(raw_prob):
    return tables.normalised(raw_prob)


@pytest.fixture(scope="session")
def table_set(prob):
    T = len(prob)
    return {
        "real": prob,
        "uniform": tables.uniform(T),
        "rowid": np.arange(1, T + 1, dtype=np.float64),
        "rowsq": ((np.arange(1, T + 1, dtype=np.int64) ** 2) % 1009 + 1).astype(np.float64),
    }


@pytest.fixture(scope="session")
def oracle():
    from oracle import loader
    loader.build()
    return loader


@pytest.fixture(scope="session")
def product_lib():
    """Path of the nvcc-built library (built on demand; nvcc cross-compiles without a GPU)."""
    from genomeassembler_dev_b200 import breakscore
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "genomeassembler_dev_b200", "csrc")], check=True)
    return breakscore.DEFAULT_LIB


@pytest.fixture(scope="session")
def emul_lib():
    """The kernel sources compiled for the CPU emulation of tests/emul/cuda_emul.h (test-only:
    checks the DEVICE ALGORITHM against the oracle on machines without a GPU)."""
    out_dir = os.path.join(ROOT, "tests", "emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libbreakscore_emul.so")
    csrc = os.path.join(ROOT, "genomeassembler_dev_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "tests", "emul", "cuda_emul.h"),
                                                                os.path.join(ROOT, "include", "breakscore.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++20", "-DBS_CPU_EMUL", "-fPIC", "-shared", "-pthread",
                        "-I", os.path.join(ROOT, "tests", "emul"), "-I", csrc, "-x", "c++",
                        os.path.join(csrc, "bs_api.cu"), os.path.join(csrc, "bs_assemble.cpp"), "-o", out], check=True)
    return out


@pytest.fixture(scope="session")
def emul_scorer(emul_lib):
    from genomeassembler_dev_b200 import breakscore
    sc = breakscore.BreakageScorer(0, emul_lib)
    yield sc
    sc.close()


@pytest.fixture(scope="session")
def gpu_scorer(product_lib):
    from genomeassembler_dev_b200 import breakscore
    sc = breakscore.BreakageScorer(0, product_lib)  # raises without a B200: no fallback
    yield sc
    sc.close()


def load_ref_vectors():
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    z = np.load(path)
    cases = []
    for i, name in enumerate(z["names"]):
        def strings(key):
            ch, off = z[f"{i}_{key}_chars"], z[f"{i}_{key}_off"]
            return [ch[off[j]:off[j + 1]].tobytes() for j in range(len(off) - 1)]
        cases.append(dict(
            name=str(name), path=strings("path"), reads=strings("reads"), truth=z[f"{i}_truth"].tobytes(),
            kmer=int(z[f"{i}_kmer"]), table=str(z[f"{i}_table"]),
            expected={k: z[f"{i}_{k}"] for k in ("sequence_len", "bp_score", "bp_score_norm_by_break_freqs",
                                                  "bp_score_norm_by_len", "kmer_breaks", "path_prob_dist_startpos",
                                                  "path_prob_dist")}))
    return cases


@pytest.fixture(scope="session")
def ref_vectors():
    return load_ref_vectors()


INT_KEYS = ("sequence_len", "kmer_breaks", "path_prob_dist_startpos")
F64_KEYS = ("bp_score", "bp_score_norm_by_break_freqs", "bp_score_norm_by_len")
KS_ATOL = 1e-12
RTOL = 1e-9  # north_star: scores and KS statistics within 1e-9 relative (fp64 accumulation)


def assert_matches_reference(got, expected, exact_scores=False):
    """got: dict from the oracle / the CUDA path; expected: arrays of the upstream list."""
    for k in INT_KEYS:
        assert np.array_equal(np.asarray(got[k]), expected[k]), (k, got[k], expected[k])
    for k in F64_KEYS:
        if exact_scores:
            assert np.array_equal(np.asarray(got[k]), expected[k]), (k, got[k], expected[k])
        else:
            np.testing.assert_allclose(got[k], expected[k], rtol=RTOL, atol=0, err_msg=k)
    pd = np.concatenate(got["path_prob_dist"]) if len(got["path_prob_dist"]) else np.zeros(0)
    assert np.array_equal(pd, expected["path_prob_dist"]), "path_prob_dist"


def assert_same_as_oracle(got, want, check_pos=True, check_hist=True):
    """CUDA path vs oracle: integer outputs bit-exact, fp64 within RTOL (NaN == NaN)."""
    for k in INT_KEYS:
        assert np.array_equal(got[k], want[k]), (k, got[k], want[k])
    if check_pos:
        assert np.array_equal(got["pos"], want["pos"]), "pos"
    if check_hist:
        assert np.array_equal(got["hist"], want["hist"]), "hist"
    for k in F64_KEYS:
        np.testing.assert_allclose(got[k], want[k], rtol=RTOL, atol=0, equal_nan=True, err_msg=k)
    # KS: R (and the oracle) accumulate +1/n.x, -1/n.y in floating point, so an exactly-zero
    # distance comes out as ~n*eps there; the device evaluates the exact ratio form.  Hence the
    # absolute floor KS_ATOL on a statistic that lives in [0, 1].
    for k in ("ks_stat_prob_dist", "ks_stat_path_freq"):
        if k in got and k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=RTOL, atol=KS_ATOL, equal_nan=True, err_msg=k)
    if "lev_dist_vs_true" in want:  # only when the caller asked the oracle for it
        assert np.array_equal(got["lev_dist_vs_true"], want["lev_dist_vs_true"]), ("lev_dist_vs_true", got["lev_dist_vs_true"], want["lev_dist_vs_true"])
    if "path_prob_dist" in got and "path_prob_dist" in want:
        for a, b in zip(got["path_prob_dist"], want["path_prob_dist"]):
            assert np.array_equal(a, b), "path_prob_dist"
