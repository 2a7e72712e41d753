"""Replay the noise-free reference-graph fixtures through the REAL reference on TensorFlow 1.15.

TEST INFRASTRUCTURE, and the one piece of it that could not be run where it was written: the
build container has no TensorFlow 1.x.  It closes the gap DESIGN §2 (ii) names.
The fixtures ``tests/golden/reference_graph_*_deterministic.npz`` were recorded from the
reference's own graph code over an eager stand-in for TF / TFP (``oracle/tf1_standin.py``), so
they pin the graph composition but restate what happens inside ``fully_connected``, fused
``batch_norm``, ``AdamOptimizer``, ``NegativeBinomial.log_prob`` ...  These cases feed
``use_deterministic_z=True`` (z = q_z_mean), with and without ``is_training``: nothing random is
left in the graph, so a machine that has

    python 3.6 / 3.7,  tensorflow>=1.15.2,<2,  tensorflow-probability==0.7,  scvae (the reference)

can run the unmodified reference on the recorded variables and minibatch and compare every
recorded tensor, the gradients of ``-lower_bound_weighted`` and the variables after one
``session.run(optimiser)`` -- TensorFlow's own arithmetic against the fixtures (fp32 tolerance).

    python oracle/check_with_tensorflow.py [fixture.npz ...]

Exit status 0 = every comparison within tolerance.  Nothing in ``tests/``, ``bench.py`` or the
product imports this file.
"""

import glob
import json
import os
import sys

import numpy

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests",
                      "golden")
RTOL = 2e-4      # the reference computes in fp32, the fixtures are fp64


def load(path):
    data = numpy.load(path)
    meta = json.loads(str(data["meta"]))
    groups = {}
    for key in data.files:
        if key != "meta":
            head, _, rest = key.partition("/")
            if head == "in":
                head, _, rest = rest.partition("/")
                head = "in_" + head
            groups.setdefault(head, {})[rest] = data[key]
    return meta, groups


def close(got, want, what, failures):
    got = numpy.asarray(got, dtype=numpy.float64).reshape(-1)
    want = numpy.asarray(want, dtype=numpy.float64).reshape(-1)
    scale = max(float(numpy.abs(want).max()), 1e-3)
    error = float(numpy.abs(got - want).max())
    ok = error <= RTOL * scale
    print("  {:<52s} max|diff| {:.2e} (scale {:.2e}) {}".format(what, error, scale,
                                                              "ok" if ok else "MISMATCH"))
    if not ok:
        failures.append(what)


def check(path):
    import tensorflow as tf
    from scvae.models import (GaussianMixtureVariationalAutoencoder, VariationalAutoencoder)
    meta, groups = load(path)
    assert meta["use_deterministic_z"], "only noise-free fixtures can be replayed"
    cls = VariationalAutoencoder if meta["model"] == "VAE" else \
        GaussianMixtureVariationalAutoencoder
    feeds = groups["in_feed"]
    model = cls(feature_size=meta["G"], log_directory="/tmp/scvae_check", **meta["kwargs"])
    failures = []
    with model.graph.as_default(), tf.Session(graph=model.graph) as session:
        session.run(tf.global_variables_initializer())
        by_name = {v.name[:-2]: v for v in tf.global_variables()}
        recorded = [name for name, _, _ in meta["variables"]]
        missing = sorted(set(recorded) - set(by_name))
        assert not missing, "variables of the fixture missing in the TF graph: {}".format(missing)
        for name in recorded:
            by_name[name].load(groups["in_var"][name].astype(numpy.float32), session)
        feed_dict = {
            model.x: feeds["X"].astype(numpy.float32), model.t: feeds["T"].astype(numpy.float32),
            model.is_training: bool(feeds["is_training"]),
            model.use_deterministic_z: True,
            model.learning_rate: float(feeds["learning_rate"]),
            model.warm_up_weight: float(feeds["warm_up_weight"]),
            model.number_of_iw_samples: int(feeds["number_of_iw_samples"]),
            model.number_of_mc_samples: int(feeds["number_of_mc_samples"]),
        }
        names = sorted(groups["out"])
        values = session.run([getattr(model, n) for n in names], feed_dict=feed_dict)
        for name, value in zip(names, values):
            close(value, groups["out"][name], name, failures)
        if meta["is_training"]:
            trainable = [n for n, _, t in meta["variables"] if t]
            gradients = tf.gradients(-model.lower_bound_weighted, [by_name[n] for n in trainable])
            live = [(n, g) for n, g in zip(trainable, gradients) if g is not None]
            for (name, _), value in zip(live, session.run([g for _, g in live], feed_dict)):
                close(value, groups["grad"][name], "grad " + name, failures)
            session.run(model.optimiser, feed_dict=feed_dict)      # BN updates, clip, Adam
            gmax = max(float(numpy.abs(g).max()) for g in groups["grad"].values())
            for name, want in sorted(groups["new"].items()):
                got = session.run(by_name[name])
                if name in groups["grad"]:     # Adam turns an fp32-noise gradient into +-lr
                    live_entries = numpy.abs(groups["grad"][name]) > 1e-3 * gmax
                    got = numpy.where(live_entries, got, want)
                close(got, want, "new " + name, failures)
    return failures


def main():
    paths = sys.argv[1:] or sorted(glob.glob(os.path.join(
        GOLDEN, "reference_graph_*deterministic*.npz")))
    failed = {}
    for path in paths:
        print(os.path.basename(path))
        failures = check(path)
        if failures:
            failed[os.path.basename(path)] = failures
    print("MISMATCHES: {}".format(json.dumps(failed, indent=1)) if failed
          else "all fixtures reproduced by TensorFlow")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
