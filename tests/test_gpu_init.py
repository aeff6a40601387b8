"""k-means++ seeding and StandardScaler on the device against sklearn-generated golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("k", [25, 60, 200])
def test_kmeans_plusplus_device_matches_sklearn(k):
    import gdr
    from gdr._dev import padded_rows
    from gdr.kmeans_init import kmeans_plusplus_device
    g = golden("kmeans_plusplus.npz")
    X = padded_rows(torch.from_numpy(g["x"]).to(DEV))
    c, idx = kmeans_plusplus_device(X, k, np.random.RandomState(int(g[f"k{k}_seed"])), return_indices=True)
    assert np.array_equal(idx.cpu().numpy(), g[f"k{k}_indices"])      # same RNG stream, same picks
    assert np.array_equal(c.cpu().numpy(), g[f"k{k}_centers"])


def test_kmeans_default_init_runs_and_is_good(oracle):
    """KMeans() with the reference's defaults (init='k-means++'): quality comparable to sklearn's."""
    import gdr
    from sklearn.cluster import KMeans as SkKMeans
    from gdr import synth
    X = synth.clustered_features(20000, 16, 40, seed=3)
    km = gdr.KMeans(n_clusters=40, random_state=0).fit(X)
    sk = SkKMeans(n_clusters=40, random_state=0, n_init=1).fit(X)
    assert km.labels_.shape == (20000,) and km.cluster_centers_.shape == (40, 16)
    assert km.inertia_ <= 1.05 * sk.inertia_
    # a wide-feature case goes through the large-shared-memory path of the scoring kernel
    Xw = synth.clustered_features(3000, 1433, 10, seed=4)
    kw = gdr.KMeans(n_clusters=140, random_state=1, max_iter=5).fit(Xw)
    assert np.isfinite(kw.inertia_)


def test_standard_scale_bit_exact():
    import gdr
    g = golden("standard_scaler.npz")
    out = gdr.standard_scale(torch.from_numpy(g["x"]).to(DEV)).cpu().numpy()
    assert np.array_equal(out, g["out"])


def test_kmeans_cluster_wrapper():
    import gdr
    from gdr import synth
    X = synth.clustered_features(5000, 64, 30, seed=9)
    labels, centers = gdr.kmeans_cluster(X, 50, seed=42)
    assert labels.dtype == np.int64 and labels.shape == (5000,) and centers.shape == (50, 64) and centers.dtype == np.float32
    assert labels.min() >= 0 and labels.max() < 50
    labels2, centers2 = gdr.kmeans_cluster(X, 50, seed=42)
    assert np.array_equal(labels, labels2) and np.array_equal(centers, centers2)   # deterministic
    lab_cap, cen_cap = gdr.kmeans_cluster(X[:20], 50, seed=1)                        # K capped to N
    assert cen_cap.shape[0] == 20
