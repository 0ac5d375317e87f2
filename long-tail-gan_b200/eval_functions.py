"""Ranking metrics with the reference's signatures (Codes/eval_functions.py:11-62), computed by the CUDA top-k kernel
(ltg_topk_metrics). `X_pred` may be a NumPy array (copied to the device; compatibility path, this is what the reference's
callers hold) or a CUDA tensor (fp32 or bf16); `heldout_batch` is a scipy CSR matrix as in the reference.
Ties are broken by lowest item index (the reference's argpartition/argsort order on ties is unspecified)."""
import numpy as np
import torch

from . import ops
from .engine import metrics_from_counts


MAX_K = 128


def _run(X_pred, heldout_batch, k, recall_ks):
    ops.init()
    if isinstance(X_pred, np.ndarray):
        scores = torch.as_tensor(np.ascontiguousarray(X_pred, dtype=np.float32)).cuda()
    else:
        scores = X_pred if X_pred.dtype in (torch.float32, torch.bfloat16) else X_pred.float()
        scores = scores.contiguous()
    n, n_items = scores.shape
    held = heldout_batch.tocsr()
    held.sort_indices()
    hp = torch.as_tensor(held.indptr.astype(np.int32)).cuda()
    hi = torch.as_tensor((held.indices if held.nnz else np.zeros(1)).astype(np.int32)).cuda()
    dcg = torch.zeros(n, dtype=torch.float64, device="cuda")
    hits = torch.zeros(n, max(1, len(recall_ks)), dtype=torch.int32, device="cuda")
    if k > MAX_K:
        raise ValueError("k = %d: the device top-k kernel ranks at most %d items per user (the reference calls these functions with "
                         "k = 100, 20 and 50: train.py:342-346, test.py:151-157)" % (k, MAX_K))
    kk = k
    ops.topk_metrics(scores, n, n_items, None, None, hp, hi, kk, recall_ks, None, dcg, hits)
    torch.cuda.synchronize()
    return metrics_from_counts(dcg.cpu().numpy(), hits.cpu().numpy(), np.diff(held.indptr.astype(np.int64)), kk, recall_ks)


def NDCG_binary_at_k_batch(X_pred, heldout_batch, k=100):
    """eval_functions.py:11-38: list of NDCG@k over the users with a non-empty held-out set."""
    return _run(X_pred, heldout_batch, k, [])["ndcg@%d" % k]


def Recall_at_k_batch(X_pred, heldout_batch, k=100):
    """eval_functions.py:40-62: (list of Recall@k over users with a non-empty held-out set, [])."""
    return _run(X_pred, heldout_batch, max(k, 1), [k])["recall@%d" % k], []
