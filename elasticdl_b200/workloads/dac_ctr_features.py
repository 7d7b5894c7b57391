"""Feature configuration of the dac_ctr model zoo entry (data of model_zoo/dac_ctr/feature_config.py:14-195,
used by feature_transform.py:36-118) and its fused GPU transform.

Raw Criteo-shaped input: 13 integer columns (tf.io.FixedLenFeature((1,), tf.int64),
elasticdl_train.py:64-76) and 26 categorical strings (8 lower-case hex characters, possibly empty).
FEATURE_GROUPS holds one feature per group; I4 is standardised but not bucketised."""
import numpy as np
import torch

from elasticdl_b200.preprocessing.layers import FeatureTransform, encode_strings

STANDARDIZED_FEATURES = ["I%d" % i for i in range(1, 14)]
FEATURES_AVGS = dict(zip(STANDARDIZED_FEATURES, [
    1.913844818114358, 105.85781137082337, 21.179428578076866, 5.735273873448716, 18067.71807784242,
    90.08603360120591, 15.626512199091756, 12.509966404126569, 101.53250047174322, 0.3374528968790535,
    2.614521353031052, 0.23277149534177055, 6.436560081179827]))
FEATURES_STDDEVS = dict(zip(STANDARDIZED_FEATURES, [
    7.203044443387521, 391.73147156506417, 354.59360229869503, 8.351369642571008, 68611.11705989522,
    340.20415627271075, 64.82617180501207, 16.71389239615237, 216.67850042198575, 0.5918310609867024,
    5.115695237395591, 2.7609291491203973, 14.799688705863462]))
FEATURE_BOUNDARIES = {
    "I1": [0.0, 1.0, 2.0, 5.0],
    "I2": [-1.0, 0.0, 1.0, 1.0, 3.0, 8.0, 23.0, 56.0, 184.0],
    "I3": [0.0, 1.0, 2.0, 4.0, 6.0, 10.0, 17.0, 36.0],
    "I4": [0.0, 1.0, 2.0, 3.0, 4.0, 6.0, 9.0, 16.0],
    "I5": [5.0, 79.0, 622.0, 1408.0, 2687.0, 4363.0, 7381.0, 13433.0, 33163.0],
    "I6": [0.0, 1.0, 7.0, 16.0, 30.0, 54.0, 98.0, 216.0],
    "I7": [0.0, 1.0, 2.0, 3.0, 5.0, 8.0, 15.0, 32.0],
    "I8": [0.0, 2.0, 3.0, 5.0, 7.0, 11.0, 16.0, 23.0, 34.0],
    "I9": [1.0, 5.0, 12.0, 21.0, 35.0, 54.0, 82.0, 134.0, 255.0],
    "I10": [0.0, 1.0],
    "I11": [0.0, 1.0, 2.0, 3.0, 6.0],
    "I12": [0.0],
    "I13": [0.0, 1.0, 2.0, 3.0, 4.0, 6.0, 10.0, 18.0],
}
# FEATURE_GROUPS (feature_config.py:156-195): I4 is absent
BUCKET_GROUP_FEATURES = ["I1", "I2", "I3", "I5", "I6", "I7", "I8", "I9", "I10", "I11", "I12", "I13"]
FEATURE_DISTINCT_COUNT = [1460, 582, 9264260, 2046299, 305, 24, 12506, 633, 3, 91211, 5670, 7659856, 3194, 27, 14876,
                          5031503, 10, 5624, 2171, 4, 6477624, 18, 15, 272811, 105, 138075]  # C1..C26
MAX_HASHING_BUCKET_SIZE = 1000000  # feature_transform.py:33
HASH_BINS = [min(c, MAX_HASHING_BUCKET_SIZE) for c in FEATURE_DISTINCT_COUNT]
assert [len(FEATURE_BOUNDARIES[f]) + 1 for f in BUCKET_GROUP_FEATURES] + HASH_BINS == \
    [5, 10, 9, 10, 9, 9, 10, 10, 3, 6, 2, 9] + HASH_BINS  # == workloads.deepfm.GROUP_ROWS


def dac_ctr_transform(ids_dtype=torch.int64):
    """transform_feature(inputs, FEATURE_GROUPS): 12 Discretization groups + 26 Hashing groups (one feature
    per group, so every ConcatenateWithOffset offset is 0) and the 13 Normalizer columns."""
    groups = [{"kind": "discretize", "column": STANDARDIZED_FEATURES.index(f), "bins": FEATURE_BOUNDARIES[f]}
              for f in BUCKET_GROUP_FEATURES]
    groups += [{"kind": "hash", "column": j, "num_bins": b} for j, b in enumerate(HASH_BINS)]
    dense = [(i, FEATURES_AVGS[f], FEATURES_STDDEVS[f]) for i, f in enumerate(STANDARDIZED_FEATURES)]
    return FeatureTransform(groups, dense, ids_dtype)


def synthetic_raw_batch(batch, seed, device, zipf_s=1.05):
    """Criteo-shaped raw features: numeric int64 [13, B] (log-normal counts, a few negatives like the
    real I2), strings uint8 [26, B, 8] (8 hex characters of a Zipf-distributed category index; ~3 %
    empty).  Returns (numeric, strings, raw string lists for the oracle)."""
    rng = np.random.RandomState(seed)
    scale = np.array([FEATURES_AVGS[f] for f in STANDARDIZED_FEATURES])[:, None]
    numeric = np.floor(rng.lognormal(0.0, 1.5, size=(13, batch)) * scale / 3.0).astype(np.int64)
    numeric[1] -= (rng.rand(batch) < 0.1) * 2  # I2 holds -1 / -2 in the real data
    raw, mats = [], []
    for j, n in enumerate(FEATURE_DISTINCT_COUNT):
        u = rng.rand(batch)
        a = 1.0 - zipf_s
        rank = np.floor(((n ** a - 1.0) * u + 1.0) ** (1.0 / a)).astype(np.int64)
        vals = (rank * 2654435761 + j * 40503) & 0xFFFFFFFF  # spread the ranks over 32-bit "hex ids"
        strs = ["%08x" % v for v in vals]
        for i in np.nonzero(rng.rand(batch) < 0.03)[0]:
            strs[i] = ""
        raw.append(strs)
        mats.append(encode_strings(strs, 8, device="cpu"))
    return torch.from_numpy(numeric).to(device), torch.stack(mats).to(device), raw
