def masked_mean(x, mask, dim=None):
    """torchoutil 0.3 ``masked_mean`` (call site: reference nn/loss/ce_mean.py:33-36; the package source is not under
    /root/reference): mean of ``x`` over ``dim`` counting only positions where ``mask`` is True; the count is clamped to
    >= 1 so an all-masked row yields 0 (restated from the published package; rows with at least one target -- every
    caption the path scores -- do not depend on the clamp)."""
    if dim is None:
        dim = ()
    return (x * mask).sum(dim=dim) / mask.sum(dim=dim).clamp(min=1.0)
