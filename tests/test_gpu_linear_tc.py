"""GPU tests of the tcgen05 GEMM next to the aggregation (csrc/linear_tc.cu, gglb200_linear_tf32x3): fp32 accuracy of the
3xTF32 split against an fp64 matmul (1e-5 relative is the north_star bar for fp32 features; plain TF32 would sit at ~5e-4),
bias / ReLU epilogue, ragged M, every K / N class the dispatcher has, autograd, and GCNConv(aggregate_first) end to end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('M,K,N', [(1000, 100, 256), (333, 128, 64), (128, 4, 4), (5000, 36, 100), (129, 64, 192), (77, 8, 16),
                                   (4096, 104, 128), (1, 100, 256),
                                   (700, 256, 64), (513, 200, 128), (300, 132, 256),      # K > 128: three / four K passes
                                   (1000, 256, 47), (77, 8, 3), (260, 100, 130)])         # N not a multiple of 4 (class counts)
@pytest.mark.parametrize('bias,relu', [(False, False), (True, True)])
def test_linear_tc_matches_fp64(M, K, N, bias, relu):
    from gammagl_b200 import mpops
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g)
    a[::3] *= 37.0                                        # mixed magnitudes: the split must hold across exponents
    w = torch.randn(K, N, generator=g) * 0.3
    b = torch.randn(N, generator=g) if bias else None
    ref = a.double() @ w.double()
    if bias:
        ref = ref + b.double()
    if relu:
        ref = torch.relu(ref)
    out = mpops.linear_tc(a.to(DEV), w.to(DEV), None if b is None else b.to(DEV), relu)
    assert out.shape == (M, N) and out.dtype == torch.float32
    assert rel(out, ref) < 1e-5
    # and it is as accurate as the fp32 library GEMM itself (not merely within the bar)
    lib = a.to(DEV) @ w.to(DEV)
    if bias:
        lib = lib + b.to(DEV)
    if relu:
        lib = torch.relu(lib)
    assert rel(out, ref) < max(8 * rel(lib, ref), 5e-6)


def test_linear_tc_autograd():
    from gammagl_b200 import mpops
    g = torch.Generator().manual_seed(5)
    M, K, N = 700, 100, 64
    a = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) * 0.2
    b = torch.randn(1, N, generator=g)
    go = torch.randn(M, N, generator=g)
    ad, wd, bd = (t.double().requires_grad_(True) for t in (a, w, b))
    torch.relu(ad @ wd + bd).backward(go.double())
    ac, wc, bc = (t.to(DEV).requires_grad_(True) for t in (a, w, b))
    mpops.linear_tc(ac, wc, bc, True).backward(go.to(DEV))
    assert rel(ac.grad, ad.grad) < 1e-5 and rel(wc.grad, wd.grad) < 1e-5 and rel(bc.grad, bd.grad) < 1e-5


def test_gcn_conv_aggregate_first_runs_the_tensor_core_gemm():
    from gammagl_b200 import _lib, mpops
    from gammagl_b200.layers import GCNConv
    N, E, F, H = 3000, 40000, 100, 256
    g = torch.Generator().manual_seed(9)
    ei = torch.randint(0, N, (2, E), generator=g)
    x = torch.randn(N, F, generator=g)
    layer = GCNConv(F, H, aggregate_first=True, fused_activation='relu').to(DEV)
    taken = []
    real = mpops.linear_tc
    mpops.linear_tc = lambda *a, **k: (taken.append(1), real(*a, **k))[1]
    try:
        xx = x.to(DEV).requires_grad_(True)
        out = layer(xx, ei.to(DEV))
        out.square().sum().backward()
    finally:
        mpops.linear_tc = real
    assert taken
    # dense fp64 restatement of gcn_conv.py:78-106 (norm='both', unit weights) + ReLU
    src, dst = ei
    A = torch.zeros(N, N, dtype=torch.float64)
    A.index_put_((dst, src), torch.ones(E, dtype=torch.float64), accumulate=True)
    ds = torch.bincount(src, minlength=N).double().pow(-0.5)
    dd = torch.bincount(dst, minlength=N).double().pow(-0.5)
    ds[torch.isinf(ds)] = 0
    dd[torch.isinf(dd)] = 0
    xd = x.double().requires_grad_(True)
    W = layer.linear.weights.detach().double().cpu()
    od = torch.relu((dd[:, None] * (A @ (ds[:, None] * xd))) @ W + layer.bias.detach().double().cpu())
    od.square().sum().backward()
    assert rel(out, od) < 1e-5
    assert rel(xx.grad, xd.grad) < 1e-5


def test_layer_linear_on_tensor_cores_and_its_input_gradient():
    """layers.Linear(tensor_core=True): forward and input gradient on the tcgen05 kernel (GGLB200_LINEAR_BWD=tc), weight / bias
    gradients from the library GEMM -- all within 1e-5 of fp64; 47 output classes run the scalar-store epilogue, their input
    gradient (K = 47, not a multiple of 4) falls back to torch.matmul."""
    from gammagl_b200.layers import Linear
    g = torch.Generator().manual_seed(2)
    M = 5000
    for fin, fout in ((256, 100), (128, 128), (256, 47)):
        lin = Linear(fin, fout, bias=True, tensor_core=True).to(DEV)
        with torch.no_grad():
            lin.biases.copy_(torch.randn(fout, generator=g))
        x = torch.randn(M, fin, generator=g)
        go = torch.randn(M, fout, generator=g)
        xx = x.to(DEV).requires_grad_(True)
        out = lin(xx)
        out.backward(go.to(DEV))
        xd = x.double().requires_grad_(True)
        W, b = lin.weights.detach().double().cpu().requires_grad_(True), lin.biases.detach().double().cpu().requires_grad_(True)
        od = xd @ W + b
        od.backward(go.double())
        assert rel(out, od) < 1e-5 and rel(xx.grad, xd.grad) < 1e-5
        assert rel(lin.weights.grad, W.grad) < 1e-5 and rel(lin.biases.grad, b.grad) < 1e-5


@pytest.mark.parametrize('env', [{'GGLB200_LINEAR': 'tile'}, {'GGLB200_LINEAR_STAGE': '0'}, {'GGLB200_LINEAR_PLAN': 'deep'},
                                 {'GGLB200_LINEAR_RING': '4', 'GGLB200_LINEAR_STAGE': '0'}],
                         ids=['one_tile_per_cta', 'per_thread_a_loads', 'plan_32_column_passes', 'ring4'])
@pytest.mark.parametrize('M,K,N', [(1000, 100, 256), (513, 200, 128), (129, 64, 192), (77, 8, 16)])
def test_linear_tc_alternative_forms(M, K, N, env, monkeypatch):
    """The switches of the tcgen05 Linear (read per call by the library): one CTA per tile, per-thread A loads instead of the staged
    tile, the 32-column plan, a deeper W ring -- every form must give the same answer as the default to the last bit of the 3xTF32
    scheme (same products, same accumulation order per output element) and stay within 1e-5 of fp64."""
    from gammagl_b200 import mpops
    g = torch.Generator().manual_seed(M * 7 + K + N)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(K, N, generator=g) * 0.3).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    ref = torch.relu(a.double() @ w.double() + b.double())
    base = mpops.linear_tc(a, w, b, True)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = mpops.linear_tc(a, w, b, True)
    assert rel(alt, ref) < 1e-5 and rel(base, ref) < 1e-5
    if 'GGLB200_LINEAR_PLAN' not in env:          # the 32-column plan splits K differently: other partial-sum order
        assert torch.equal(alt, base)
