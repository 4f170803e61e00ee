"""
NOT COLLECTED (no test_ prefix): GPU checks written after the round's GPU budget was spent -- a cold box charged 10.5
minutes for one 3-minute test pass -- so they have never run on a B200.  Run with
    python -m pytest tests/pending_gpu_checks.py -q -m gpu
and move them into tests/test_gpu_estimators.py once green.

What they cover: `KerasRawModelRegressor` fit / predict inside a Pipeline (the reference's
tests/gordo/machine/model/test_raw_keras.py::test_raw_keras_part_of_pipeline) and, with it, `estimator.predict(X)` for a
model whose n_features_out differs from n_features -- which raised until `FFFleet.score` learned to supply a zero target
for pure forward passes (the C-ABI refuses a NULL y when the widths differ).
"""
import numpy as np
import pytest

from oracle import dense

pytestmark = pytest.mark.gpu


def test_raw_keras_regressor_in_a_pipeline_fits_and_matches_the_oracle():
    """The reference's test_raw_keras_part_of_pipeline (PCA -> KerasRawModelRegressor: Dense(4) -> Dense(1), y != X), then
    the fitted weights' forward pass against the oracle's Dense forward."""
    import yaml
    from gordo_b200 import serializer
    rng = np.random.default_rng(0)
    X, y = rng.random((100, 4)), rng.random((100, 1))
    config = yaml.safe_load("""
    sklearn.pipeline.Pipeline:
        steps:
            - sklearn.decomposition.PCA:
                n_components: 4
            - gordo.machine.model.models.KerasRawModelRegressor:
                kind:
                    compile:
                        loss: mse
                        optimizer: adam
                    spec:
                        tensorflow.keras.models.Sequential:
                            layers:
                                - tensorflow.keras.layers.Dense:
                                    units: 4
                                    input_shape: [4]
                                    activation: tanh
                                - tensorflow.keras.layers.Dense:
                                    units: 1
                epochs: 5
    """)
    pipe = serializer.from_definition(config, redirect_gordo=True)
    pipe.fit(X, y)
    out = pipe.predict(X)
    assert out.shape == (100, 1) and np.isfinite(out).all()
    est = pipe.steps[1][1]
    hist = est.get_metadata()["history"]
    assert len(hist["loss"]) == 5
    spec = {"type": "ff", "widths": [4, 4, 1], "acts": ["tanh", "linear"], "l1": [0.0, 0.0]}
    Z = pipe.steps[0][1].transform(X).astype(np.float32)
    want = dense.ff_forward(spec, dense.ff_unflatten(est.model.params, spec["widths"]), Z)
    np.testing.assert_allclose(out, want, rtol=0, atol=2e-5)
