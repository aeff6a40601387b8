"""k-means++ seeding on the device (sklearn/_kmeans.py:180-278).  SURVEY §8(f) item 2 —
a "next" row: until it lands the estimator must be given ``init=<array>`` or
``init="random"``; asking for k-means++ fails loudly instead of falling back to the CPU."""


def kmeans_plusplus_device(Xc, n_clusters, random_state):
    raise NotImplementedError(
        "init='k-means++' is not implemented on the device yet; pass init=<array> "
        "(parity mode) or init='random'")
