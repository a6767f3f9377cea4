"""The only arithmetic of the reference's ``scripts/search.py`` that belongs to the hot path: the FLOPS relevance/cost
metric computed from the per-token document frequencies that ``SparseEncoder`` accumulates (reference :82-90).
Query encoding itself is ``SparseEncoder.encode(texts, inf_free=...)``; the OpenSearch round trip is out of scope."""
import torch


def flops_metric(count_q, num_queries, count_d, num_docs):
    """Expected number of multiply-adds per query-document pair: sum_v P(v in query) * P(v in doc).

    Returns (flops, avg query length, avg doc length) like the reference's search() bookkeeping."""
    pq = count_q.to(torch.float32) / float(num_queries)
    pd = count_d.to(torch.float32) / float(num_docs)
    return float((pq * pd).sum()), float(pq.sum()), float(pd.sum())
