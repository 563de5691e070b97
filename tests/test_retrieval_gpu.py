"""GPU: retrieval top-k (classifier.py:963-983) on fp32 features. Indices must be bit-exact against the
float64 evaluation of the reference formula, reproduce the golden vectors generated from the real
reference, and agree with the fp32 oracle everywhere except the oracle's own fp32 near-ties."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda:0"
# Rows (of 3783) whose top-50 / top-1 lists differ from the reference's own fp32 path (cuBLAS sgemm + torch.topk) on the
# seed-7 features: measured value + 1 (profiles/r02_measured_parity.jsonl). Every such row is an fp32 near-tie of the
# reference (asserted below: the swapped similarities differ by < 1e-6); against the fp64 evaluation of the reference
# formula the indices are bit-exact for all 3783 rows.
MAX_ROWS_DIFF_TOP50 = 6      # measured: 5 of 3783 rows
MAX_ROWS_DIFF_TOP1 = 1       # measured: 0


def _f64_reference(te, tr, kmax):
    te, tr = te.double(), tr.double()
    te = te - te.mean(0, keepdim=True)
    tr = tr - tr.mean(0, keepdim=True)
    te = te / te.norm(dim=1, keepdim=True).clamp_min(1e-12)
    tr = tr / tr.norm(dim=1, keepdim=True).clamp_min(1e-12)
    sim = te @ tr.t()
    return sim, torch.topk(sim, kmax, dim=1)[1]


def test_retrieval_matches_golden_reference(golden_dir):
    from dualvar_b200.retrieval import retrieval_topk
    g = np.load(os.path.join(golden_dir, "objectives.npz"))
    te, tr = torch.from_numpy(g["ret_test"]).to(dev), torch.from_numpy(g["ret_train"]).to(dev)
    sim, idx = retrieval_topk(te, tr)
    np.testing.assert_allclose(sim.cpu().numpy(), g["ret_sim"], rtol=1e-5, atol=1e-6)
    for k in (1, 5, 10, 20, 50):
        assert np.array_equal(idx[k].cpu().numpy(), g[f"ret_top{k}"])


def test_retrieval_ucf101_shape_bit_exact_indices(measured):
    """Config 5 shape: 3783 test x 9537 train x 512-d features (UCF101 split 1), seed 7."""
    from dualvar_b200.retrieval import retrieval_topk, retrieval_accuracy
    from oracle import objectives as OO
    gen = torch.Generator(device=dev).manual_seed(7)
    te = torch.randn(3783, 512, device=dev, generator=gen)
    tr = torch.randn(9537, 512, device=dev, generator=gen)
    sim, idx = retrieval_topk(te, tr)
    sim64, ref_idx = _f64_reference(te, tr, 50)
    assert torch.equal(idx[50], ref_idx)                      # bit-exact against the fp64 formula
    for k in (1, 5, 10, 20):
        assert torch.equal(idx[k], ref_idx[:, :k])
    torch.testing.assert_close(sim, sim64.float(), rtol=0, atol=2e-7)
    # fp32 oracle (the reference's own arithmetic): identical except at its fp32 near-ties
    torch.backends.cuda.matmul.allow_tf32 = False
    sim32, idx32 = OO.retrieval_topk(te, tr)
    n_rows_diff = int((idx32[50] != idx[50]).any(dim=1).sum())
    n_top1_diff = int((idx32[1] != idx[1]).sum())
    measured("retrieval.rows_differing_from_fp32_oracle_top50", n_rows_diff)
    measured("retrieval.rows_differing_from_fp32_oracle_top1", n_top1_diff)
    # exact counts on this image's cuBLAS (sgemm summation order): see MAX_ROWS_DIFF below
    assert n_rows_diff <= MAX_ROWS_DIFF_TOP50, n_rows_diff
    assert n_top1_diff <= MAX_ROWS_DIFF_TOP1, n_top1_diff
    diff = idx32[50] != idx[50]
    if diff.any():
        # every disagreement swaps two gallery items whose reference similarities differ by < 1e-6
        a = sim32.gather(1, idx32[50])[diff]
        b = sim32.gather(1, idx[50])[diff]
        assert (a - b).abs().max().item() < 1e-6
    labels_tr = torch.randint(0, 101, (9537,), device=dev, generator=gen)
    labels_te = torch.randint(0, 101, (3783,), device=dev, generator=gen)
    acc = retrieval_accuracy(idx, labels_tr, labels_te)
    acc_ref = OO.retrieval_accuracy({k: ref_idx[:, :k] for k in (1, 5, 10, 20, 50)}, labels_tr, labels_te)
    assert acc == acc_ref and acc[1] <= acc[5] <= acc[10] <= acc[20] <= acc[50]


def test_retrieval_ties_break_to_lowest_index():
    from dualvar_b200.retrieval import retrieval_topk
    base = torch.randn(6, 32, device=dev)
    train = torch.cat([base, base, base])          # every gallery vector appears three times
    test = torch.randn(5, 32, device=dev)
    _, idx = retrieval_topk(test, train, ks=(3,))
    top = idx[3]
    assert bool((top[:, 1] == top[:, 0] + 6).all()) and bool((top[:, 2] == top[:, 0] + 12).all())
