"""monai.utils subset (0.7.0)."""


def ensure_tuple_rep(tup, dim):
    if isinstance(tup, (list, tuple)):
        if len(tup) == dim:
            return tuple(tup)
        raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")
    return (tup,) * dim
