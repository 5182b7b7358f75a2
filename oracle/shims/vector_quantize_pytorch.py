"""Stand-in for the pip package imported (but never used on the hot path) at
distilcodec/vector_quantization/grfsq.py:7. The VQ code the path uses is vendored in the reference."""


class GroupedResidualFSQ:  # noqa: D401
    pass


class GroupedResidualVQ:
    pass
