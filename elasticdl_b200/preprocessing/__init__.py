from elasticdl_b200.preprocessing.layers import (  # noqa: F401
    ConcatenateWithOffset,
    Discretization,
    FeatureTransform,
    Hashing,
    Normalizer,
    encode_strings,
)
