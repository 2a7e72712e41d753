"""The CUDA SOURCE of the small kernels written without a device (sampled KL, dropout), compiled
for the host and executed thread by thread.

``csrc/dropout.cu`` and the sampled-KL kernels of ``csrc/latent.cu`` were written after the GPU
budget of their round was spent.  Their arithmetic is restated in numpy
(``tests/test_sampled_kl_math.py``, ``tests/kernel_standins.py``) and held to autograd and to the
reference-graph fixtures -- but a restatement does not execute the kernel text.  This test does:
it cuts the ``__global__`` functions out of the ``.cu`` files, compiles them with g++ against a
few lines that stand in for the CUDA built-ins (``blockIdx`` / ``threadIdx`` as globals, every
thread of the launch grid run in sequence; ``warp_sum`` / ``block_sum`` resolved by running the
grid twice: the first pass records every thread's contribution to each reduction, the second
returns the totals; kernels without reductions run once) and compares the outputs with the
restatements.  It cannot see CUDA-specific behaviour (memory model, launch configuration limits); indexing, strides, masks
and formulas it does see.  Test infrastructure only; skipped where g++ is missing.
"""
import ctypes
import os
import re
import shutil
import subprocess

import numpy
import pytest
import torch

import kernel_standins as C
from test_sampled_kl_math import bound_rows, sampled_kl_bwd, sampled_kl_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scvae_b200", "csrc")

PRELUDE = r"""
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>
#define __global__
#define __restrict__
#define __expf expf
struct Dim3 { int x, y, z; };
static Dim3 blockIdx, threadIdx, blockDim;
constexpr int kRowsPerBlock = 4;
constexpr float kLogTiny = -87.33654475f;
// reductions across threads: pass 0 records, pass 1 answers
static int g_pass = 0;
static std::map<long long, std::vector<float>> g_table;   // (block, unit, call) -> contributions
static std::map<long long, int> g_calls;                   // per thread: reductions seen so far
static long long block_id() { return (long long)blockIdx.y * 1000003 + blockIdx.x; }
static float reduce(float v, int unit, int lane, int lanes) {
    const long long thread = block_id() * 4096 + threadIdx.x;
    const int call = g_calls[thread]++;
    const long long key = (block_id() * 64 + unit) * 4096 + call;
    if (g_pass == 0) {
        auto &slot = g_table[key];
        if ((int)slot.size() < lanes) slot.resize(lanes, 0.f);
        slot[lane] = v;
        return v;
    }
    float total = 0.f;
    for (float c : g_table[key]) total += c;
    return total;
}
static float warp_sum(float v) { return reduce(v, 1 + (threadIdx.x >> 5), threadIdx.x & 31, 32); }
static float block_sum(float v, float *) { return reduce(v, 0, threadIdx.x, blockDim.x); }
#define __shared__ static
template <typename F> static void launch(int gx, int gy, int bx, F f, int passes = 2) {
    blockDim = {bx, 1, 1};
    g_table.clear();
    for (g_pass = 0; g_pass < passes; ++g_pass) {
        g_calls.clear();
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x)
                for (int t = 0; t < bx; ++t) {
                    blockIdx = {x, y, 0};
                    threadIdx = {t, 0, 0};
                    f();
                }
    }
}
"""

ENTRY = r"""
extern "C" {
void emu_dropout_fwd(const float *x, int64_t ldx, int rows, int n, int skip, const float *noise,
                     float thr, float keep, float *out, int64_t ldo, int width) {
    launch(rows, (width + 127) / 128, 128, [&] {
        dropout_fwd_kernel(x, ldx, n, skip, noise, thr, 1.f / keep, out, ldo, width); }, 1);
}
void emu_dropout_bwd(float *dx, int64_t lddx, int rows, int n, int skip, const float *noise,
                     float thr, float keep, const float *dsrc, int64_t ldds, int acc) {
    launch(rows, (n + 127) / 128, 128, [&] {
        dropout_bwd_kernel(dx, lddx, n, skip, noise, thr, 1.f / keep, dsrc, ldds, acc); }, 1);
}
void emu_sampled_kl(const float *ph, int64_t ldph, int B, int L, int RS, const float *eps, int uv,
                    int det, float *kl_rows, float *kl_elem) {
    launch((B + 3) / 4, 1, 128, [&] {
        gaussian_sampled_kl_kernel(ph, ldph, B, L, RS, eps, uv, det, kl_rows, kl_elem); });
}
void emu_sampled_kl_bwd(const float *ph, int64_t ldph, int B, int L, int RS, const float *eps,
                        int uv, const float *dz, int64_t lddz, const float *go, float weight,
                        float coef, float *dph, int64_t lddph) {
    launch((B + 3) / 4, 1, 128, [&] {
        gaussian_sampled_kl_bwd_kernel(ph, ldph, B, L, RS, eps, uv, dz, lddz, go, weight, coef,
                                       dph, lddph); }, 1);
}
void emu_mixture_moments(const float *a, int64_t lda, const float *lse, const float *count_sum,
                         int B, int G, int RS, int K, const float *y, int64_t ldy, float *mean,
                         float *stddev, float *stddev_of_mean, int64_t ldo) {
    launch(B, (G + 255) / 256, 256, [&] {
        constrained_poisson_mixture_moments_kernel(a, lda, lse, count_sum, B, G, RS, K, y, ldy,
                                                   mean, stddev, stddev_of_mean, ldo); }, 1);
}
void emu_bound_rows(const float *logp, const float *kl_rows, int R, int S, int B, float weight,
                    float *out, float *go) {
    launch(1, 1, (S * B >= 1024) ? 1024 : 256, [&] {
        vae_bound_rows_kernel(logp, kl_rows, R, S, B, weight, out, go); });
}
}
"""


def _kernel(text, name):
    start = text.index("__global__ void", text.rfind("\n\n", 0, text.index(name + "(")))
    depth, k = 0, text.index("{", text.index(name + "("))
    while True:
        depth += {"{": 1, "}": -1}.get(text[k], 0)
        if depth == 0:
            break
        k += 1
    return re.sub(r"__launch_bounds__\([^)]*\)\s*", "", text[start:k + 1])


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    dropout = open(os.path.join(CSRC, "dropout.cu")).read()
    latent = open(os.path.join(CSRC, "latent.cu")).read()
    source = PRELUDE + "\n\n".join([
        _kernel(dropout, "dropout_fwd_kernel"), _kernel(dropout, "dropout_bwd_kernel"),
        _kernel(latent, "gaussian_sampled_kl_kernel"),
        _kernel(latent, "gaussian_sampled_kl_bwd_kernel"),
        _kernel(latent, "vae_bound_rows_kernel"),
        _kernel(open(os.path.join(CSRC, "constrained_poisson.cu")).read(),
                "constrained_poisson_mixture_moments_kernel")]) + ENTRY
    directory = tmp_path_factory.mktemp("emu")
    path = os.path.join(str(directory), "emu.cpp")
    with open(path, "w") as handle:
        handle.write(source)
    library = os.path.join(str(directory), "libemu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", library, path],
                   check=True)
    return ctypes.CDLL(library)


P, I, L64, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


def ptr(tensor):
    return P(tensor.data_ptr()) if tensor is not None else None


def test_dropout_kernel_source(emulated):
    gen = torch.Generator().manual_seed(9)
    rows, n, width, keep, thr = 37, 11, 16, 0.8, 0.8416212335729143
    x = torch.randn(rows, width, generator=gen)
    noise = torch.randn(rows, n, generator=gen)
    d = torch.randn(rows, width, generator=gen)
    acc = torch.randn(rows, width, generator=gen)
    for skip in (7, n, 0):
        want = torch.zeros(rows, width)
        C.dropout_fwd(x, rows, n, skip, noise, thr, keep, want, width)
        got = torch.zeros(rows, width)
        emulated.emu_dropout_fwd(ptr(x), L64(width), I(rows), I(n), I(skip), ptr(noise), F(thr),
                                 F(keep), ptr(got), L64(width), I(width))
        assert torch.allclose(got, want, rtol=1e-6, atol=0)
        for dsrc, accumulate in ((None, 0), (d, 0), (d, 1)):
            want = acc.clone()
            C.dropout_bwd(want, rows, n, skip, noise, thr, keep, dsrc=dsrc,
                          accumulate=bool(accumulate))
            got = acc.clone()
            emulated.emu_dropout_bwd(ptr(got), L64(width), I(rows), I(n), I(skip), ptr(noise),
                                     F(thr), F(keep), ptr(dsrc), L64(width), I(accumulate))
            assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("R,S,unit_variance", [(1, 1, 0), (3, 2, 0), (2, 1, 1)])
def test_sampled_kl_kernel_sources(emulated, R, S, unit_variance):
    gen = torch.Generator().manual_seed(3)
    B, L, RS, weight = 37, 5, R * S, 0.6
    nL = L if unit_variance else 2 * L
    ph = torch.randn(B, nL + 2, generator=gen)
    if not unit_variance:
        ph[:, L:2 * L] *= 2.5
    eps = torch.randn(RS * B, L, generator=gen)
    dz = torch.randn(RS * B, L + 3, generator=gen)
    logp = torch.randn(RS * B, generator=gen) * 5 - 100
    ph64, eps64 = ph[:, :nL].double().numpy(), eps.double().numpy()

    for deterministic in (0, 1):
        rows = torch.zeros(RS * B)
        elem = torch.zeros(B, L)
        emulated.emu_sampled_kl(ptr(ph), L64(nL + 2), I(B), I(L), I(RS),
                                None if deterministic else ptr(eps), I(unit_variance),
                                I(deterministic), ptr(rows), ptr(elem))
        ref_rows, ref_elem = sampled_kl_rows(ph64, None if deterministic else eps64, B, L, RS,
                                             bool(unit_variance), bool(deterministic))
        assert numpy.abs(rows[:ref_rows.size].numpy() - ref_rows).max() <= \
            1e-5 * numpy.abs(ref_rows).max()
        assert numpy.abs(elem.numpy() - ref_elem).max() <= 1e-5 * numpy.abs(ref_elem).max()

    kl_rows, _ = sampled_kl_rows(ph64, eps64, B, L, RS, bool(unit_variance))
    kl32 = torch.tensor(kl_rows, dtype=torch.float32)
    out, go = torch.zeros(4), torch.zeros(RS * B)
    emulated.emu_bound_rows(ptr(logp), ptr(kl32), I(R), I(S), I(B), F(weight), ptr(out), ptr(go))
    ref_out, ref_go = bound_rows(logp.double().numpy(), kl32.double().numpy(), R, S, B, weight)
    assert numpy.abs(out.numpy() - ref_out).max() <= 2e-5 * numpy.abs(ref_out).max()
    assert numpy.abs(go.numpy() - ref_go).max() <= 1e-4 * numpy.abs(ref_go).max()

    for upstream, coef in ((go, 0.0), (None, weight / (S * B))):
        dph = torch.zeros(B, nL + 2)
        emulated.emu_sampled_kl_bwd(ptr(ph), L64(nL + 2), I(B), I(L), I(RS), ptr(eps),
                                    I(unit_variance), ptr(dz), L64(L + 3), ptr(upstream),
                                    F(weight), F(coef), ptr(dph), L64(nL + 2))
        ref = sampled_kl_bwd(ph64, eps64, dz[:, :L].double().numpy(),
                             None if upstream is None else upstream.double().numpy(), weight,
                             coef, B, L, RS, bool(unit_variance))
        assert numpy.abs(dph[:, :nL].numpy() - ref).max() <= 1e-5 * numpy.abs(ref).max()
        assert float(dph[:, nL:].abs().max()) == 0.0         # nothing written past the heads


def test_constrained_poisson_mixture_moments_kernel_source(emulated):
    gen = torch.Generator().manual_seed(5)
    B, G, RS, K = 7, 300, 2, 3
    rows = K * RS * B
    a = torch.randn(rows, G + 4, generator=gen) * 2
    lse = torch.logsumexp(a[:, :G].double(), dim=1).float()
    count_sum = torch.rand(B, generator=gen) * 500 + 20
    y = torch.softmax(torch.randn(B, K + 1, generator=gen), dim=-1)       # padded leading dimension
    y[:, :K] = torch.softmax(torch.randn(B, K, generator=gen), dim=-1)
    outs = [torch.zeros(B, G + 4) for _ in range(3)]
    emulated.emu_mixture_moments(ptr(a), L64(G + 4), ptr(lse), ptr(count_sum), I(B), I(G), I(RS),
                                 I(K), ptr(y), L64(K + 1), ptr(outs[0]), ptr(outs[1]),
                                 ptr(outs[2]), L64(G + 4))
    want = [torch.zeros(B, G) for _ in range(3)]
    C.constrained_poisson_mixture_moments(a, lse, count_sum, B, G, RS, K,
                                          y[:, :K].contiguous(), *want)
    for got, ref in zip(outs, want):
        assert torch.allclose(got[:, :G], ref, rtol=2e-5, atol=1e-6)
        assert float(got[:, G:].abs().max()) == 0.0
