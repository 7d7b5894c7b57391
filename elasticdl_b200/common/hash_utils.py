"""Partition functions of the PS path (host-side integer hashing).

Mirrors elasticdl/python/common/hash_utils.py:17-23 and the Go twins
go/pkg/ps/checkpoint.go:31-44.  The row scatter of hash_utils.py:26-62 is done on
the device (id % N inside the kernels), so it has no host function here.
"""
import hashlib


def string_to_id(name, bucket_num):
    """Dense parameter -> shard.  The sha256 HEX digest is parsed in radix 32 (sic):
    both reference implementations agree on that (hash_utils.py:19, checkpoint.go:36)."""
    h = hashlib.sha256(name.encode("utf-8"))
    return int(h.hexdigest(), base=32) % bucket_num


def int_to_id(number, bucket_num):
    """Row id -> shard (hash_utils.py:22-23)."""
    return number % bucket_num
