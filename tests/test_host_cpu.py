"""Host-side logic that needs no GPU: CLI surface, option validation, model naming."""
import pytest


def test_cli_accepts_the_reference_flags_of_the_hot_path():
    from scvae_b200 import cli
    a = cli._parser().parse_args(
        ["train", "development", "-m", "GMVAE", "-K", "4", "-r", "negative binomial", "-k", "2",
         "--bc", "--count-sum", "--generative-architecture", "LFM", "-l", "8", "-H", "64", "32",
         "-e", "3", "-B", "128"])
    assert a.model_type == "GMVAE" and a.number_of_classes == 4
    assert a.number_of_reconstruction_classes == 2 and a.batch_correction and a.count_sum
    assert a.generative_architecture == "LFM" and a.hidden_sizes == [64, 32]
    e = cli._parser().parse_args(["evaluate", "development", "--model-versions", "end_of_training", "best_model"])
    assert e.func is cli.evaluate


def test_option_validation_mirrors_the_scope():
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200.gaussian_mixture_variational_autoencoder import (
        GaussianMixtureVariationalAutoencoder)
    # built this round: -k, batch correction, count sum, LFM, constrained Poisson (VAE)
    model = VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                   reconstruction_distribution="negative binomial",
                                   number_of_reconstruction_classes=3, batch_correction=True,
                                   number_of_batches=2, count_sum=True,
                                   generative_architecture="LFM")
    assert "k_3" in model.name and "bc" in model.name and "ga_LFM" in model.name
    VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                           reconstruction_distribution="constrained poisson")
    GaussianMixtureVariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                          number_of_latent_clusters=3,
                                          reconstruction_distribution="negative binomial",
                                          number_of_reconstruction_classes=2, batch_correction=True,
                                          number_of_batches=2)
    with pytest.raises(ValueError):      # as the reference (MU:883-897): no -k around zero inflation
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               reconstruction_distribution="zero-inflated poisson",
                               number_of_reconstruction_classes=2)
    # still outside: loud, never silently degraded
    VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                           dropout_keep_probabilities=[0.9])          # VAE: built
    model = GaussianMixtureVariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                                  number_of_latent_clusters=3,
                                                  dropout_keep_probabilities=[0.9, 0.8, 0.7, 0.6])
    assert "dropout_0.9_0.8_0.7_0.6" in model.name          # GMVAE: built, four keep probabilities
    GaussianMixtureVariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                          number_of_latent_clusters=3,
                                          reconstruction_distribution="constrained poisson")
    # the continuous / binary reconstruction distributions are built (csrc/continuous.cu)
    for name in ("gamma", "gaussian", "modified gaussian", "log-normal", "bernoulli", "lomax",
                 "exponentially_modified_gaussian"):
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               reconstruction_distribution=name)
    with pytest.raises(NotImplementedError):
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               reconstruction_distribution="multinomial")
    with pytest.raises(TypeError):
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               batch_correction=True)          # number of batches missing


def test_packed_stream_slabs_decode_to_the_rows_of_the_matrix():
    """Host logic of the streamed wire format (scvae_csr_densify_packed / scvae_pack_row_slab):
    every slab of an epoch -- assembled by the native host packer from the per-row strings, through
    the feeder thread -- decoded with numpy exactly as the kernel reads it, reproduces its rows in
    epoch order, escapes (counts >= 255) included."""
    import numpy
    import scipy.sparse
    from scvae_b200 import kernels as K
    from scvae_b200.hotloop import PackedStream
    rng = numpy.random.RandomState(0)
    n, G, B = 37, 700, 8
    dense = ((rng.rand(n, G) < 0.1) * rng.randint(1, 400, size=(n, G))).astype(numpy.float32)
    dense[3] = 0                      # an empty row
    dense[5, :] = 7                   # a full row: every block holds 255 entries
    dense[6, :] = 300                 # every entry an escape
    stream = PackedStream(scipy.sparse.csr_matrix(dense), "cpu", B, pack_threads=3)
    assert stream.nblk == 3
    order = rng.permutation(n)
    assert stream.pack_epoch(order) == 5
    nblk = stream.nblk
    seen = 0
    for k in range(5):
        slot = stream.fetch(k % 2, k)
        rows, host = slot["rows"], slot["host"]
        assert rows == min(B, n - k * B) and host.size == slot["bytes"]
        off = host[:4 * (rows + 1)].view(numpy.int32)
        consts = host[4 * (rows + 1):4 * (rows + 1) + 4 * rows].view(numpy.float32)
        strings = host[K.packed_rows_offset(rows):]
        for r in range(rows):
            s = strings[off[r]:off[r + 1]].astype(numpy.int64)
            nesc, nnz = s[0] | (s[1] << 8), s[2] | (s[3] << 8)
            blocks = s[4:4 + nblk]
            assert blocks.sum() == nnz and s.size % 16 == 0
            assert s.size == (4 + nblk + 2 * nnz + 4 * nesc + 15) // 16 * 16
            e = s[4 + nblk:4 + nblk + 2 * nnz].reshape(nnz, 2)
            esc = s[4 + nblk + 2 * nnz:4 + nblk + 2 * nnz + 4 * nesc].reshape(nesc, 4)
            values = e[:, 1].copy()
            lookup = {int(a | (b << 8)): int(c | (d << 8)) for a, b, c, d in esc}
            for i in numpy.nonzero(values == 255)[0]:
                values[i] = lookup[int(i)]
            out = numpy.zeros(G, numpy.float32)
            out[numpy.repeat(numpy.arange(nblk), blocks) * 255 + e[:, 0]] = values
            want = dense[order[k * B + r]]
            assert numpy.array_equal(out, want)
            assert consts[r] == stream.row_const[order[k * B + r]]
            seen += 1
    assert seen == n
    stream.close()
    stream.pack_epoch(order)
    with pytest.raises(ValueError):                      # slabs leave in epoch order
        stream.fetch(0, 1)
    stream.close()
    # ~2 bytes per non-zero when the counts fit one byte
    small = PackedStream(scipy.sparse.csr_matrix(numpy.minimum(dense, 200.0)), "cpu", B)
    assert small.bytes_per_nonzero < 2.0 + (4 + nblk + 15) * n / (dense > 0).sum() + 1e-9
