"""CPU checks of the host logic added with the bf16x3 mode and the CTA-pair GEMM (no GPU, no library call):
  * the hi / lo plane arithmetic (numpy restatement of fn_split_bf16 and of the three-plane product the tcgen05 kernels
    evaluate): representation error <= 2^-16, product error ~1e-5 of the operand scale -- the budget behind the 1e-3 bar;
  * the bf16x3 layouts the kernels assume (gate planes 128*H elements apart in the blocked layout; the plane-product ->
    column map of the state / gate-gradient loaders in csrc/fn_gru_tc.cu);
  * the split-K planner (ops_bf16.plan_splits) and the pair-kernel eligibility mirror;
  * the fp32-gradient <-> bf16-view hand-over of split activations (ops_x3.as_split_grad / from_split_grad)."""
import numpy as np
import pytest
import torch

from test_layouts import gate_off


def bf16_round(x):
    """fp32 -> bf16 (round to nearest even) -> fp32, numpy."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def split(x):
    hi = bf16_round(x)
    return hi, bf16_round(x - hi)


def test_planes_carry_16_mantissa_bits():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-20, 20, 100000))).astype(np.float32)
    hi, lo = split(x)
    assert np.array_equal(hi, torch.from_numpy(x).to(torch.bfloat16).float().numpy())      # same rounding as torch / the kernels
    rel = np.abs((hi.astype(np.float64) + lo) - x) / np.abs(x)
    assert rel.max() <= 2.0 ** -16 and np.sqrt((rel ** 2).mean()) < 2.0 ** -18


@pytest.mark.parametrize("K", [64, 1024])
def test_three_plane_product_error_budget(K):
    """(A_hi + A_lo)(B_hi + B_lo) - A_lo B_lo against the exact product: RMS error a few 1e-6 of the result's RMS -- ~500x
    better than one bf16 product, ~10x worse than fp32 FMA (tools/x3_diag.py measures the same on the GPU: 5e-6)."""
    rng = np.random.default_rng(1)
    A, B = rng.standard_normal((64, K)).astype(np.float32), rng.standard_normal((48, K)).astype(np.float32)
    ah, al = split(A); bh, bl = split(B)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    x3 = ah.astype(np.float64) @ bh.T + al.astype(np.float64) @ bh.T + ah.astype(np.float64) @ bl.T
    bf = ah.astype(np.float64) @ bh.T
    rms = np.sqrt((ref ** 2).mean())
    e3, e1 = np.sqrt(((x3 - ref) ** 2).mean()) / rms, np.sqrt(((bf - ref) ** 2).mean()) / rms
    assert e3 < 1e-5 and e1 > 100 * e3


def test_x3_gate_planes_and_loader_column_map():
    B, H = 70, 64
    # saved gates: the lo plane of column c is column c + 4H of an 8H-wide blocked row = 128*H elements further on
    for t, b, c in ((0, 0, 0), (3, 69, 4 * H - 1), (1, 33, 2 * H + 5)):
        assert gate_off(t, b, c + 4 * H, B, 8 * H) - gate_off(t, b, c, B, 8 * H) == 128 * H
    # forward loader: K' = 3H over a state row [hi H | lo H]; products (hi, lo, hi) against weights laid out [hi | hi | lo]
    def fwd_col(k):
        prod, kk = divmod(k, H)
        return kk + (H if prod == 1 else 0)
    assert [fwd_col(k) for k in (0, H - 1, H, 2 * H - 1, 2 * H, 3 * H - 1)] == [0, H - 1, H, 2 * H - 1, 0, H - 1]
    # backward loader: K' = 9H over a gate-gradient row [hi (dr dz dn dnr) | lo (...)]; the product consumes (dr, dz, dn*r)
    def bwd_col(k):
        prod, kk = divmod(k, 3 * H)
        return (kk + H if kk >= 2 * H else kk) + (4 * H if prod == 1 else 0)
    assert bwd_col(0) == 0 and bwd_col(2 * H) == 3 * H and bwd_col(3 * H) == 4 * H and bwd_col(5 * H) == 7 * H
    assert bwd_col(6 * H) == 0 and bwd_col(9 * H - 1) == 4 * H - 1
    cols = {bwd_col(k) for k in range(9 * H)}
    assert not any(2 * H <= c < 3 * H or 6 * H <= c < 7 * H for c in cols)        # dn itself never enters the recurrence


def test_split_k_planner():
    from fadernets_b200.ops_bf16 import pair_kernel_takes, plan_splits
    assert pair_kernel_takes(2048, 1024, 131072) and pair_kernel_takes(3072, 342, 131072) and pair_kernel_takes(131072, 342, 1024)
    assert not pair_kernel_takes(131072, 16, 1024) and not pair_kernel_takes(128, 1024, 1024) and not pair_kernel_takes(512, 256, 512)
    for M, N, K in ((2048, 1024, 131072), (1024, 1024, 131072), (3072, 1024, 131072), (3072, 342, 131072), (1536, 512, 16384),
                    (3072, 16, 131072), (8192, 3072, 1024), (512, 512, 640)):
        for nprod in (1, 3):
            s = plan_splits(M, N, K, nprod)
            assert 1 <= s <= 32 and (s == 1 or K * nprod // s >= 1024)
    assert plan_splits(8192, 3072, 1024) == 1                      # short K: no split
    assert plan_splits(2048, 1024, 131072) > 1                     # the T*B-row weight gradient is split
    # the pair kernel's last wave is (nearly) full for the weight-gradient shapes of config 3
    for M, N in ((2048, 1024), (1024, 1024), (3072, 1024)):
        items = (M // 256) * (N // 256) * plan_splits(M, N, 131072)
        assert items / (-(-items // 74) * 74) > 0.95


def test_split_gradient_view_round_trip():
    from fadernets_b200.ops_x3 import as_split_grad, from_split_grad
    g = torch.randn(5, 3, 64)
    v = as_split_grad(g)
    assert v.dtype == torch.bfloat16 and v.shape == (5, 3, 128)
    assert torch.equal(from_split_grad(v), g)
    assert torch.equal(from_split_grad(torch.zeros(2, 128, dtype=torch.bfloat16)), torch.zeros(2, 64))   # materialised zero gradient
