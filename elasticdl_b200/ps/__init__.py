from elasticdl_b200.ps.group import PSGroup  # noqa: F401
