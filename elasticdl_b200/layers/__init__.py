from elasticdl_b200.layers.embedding import Embedding  # noqa: F401
