"""
oracle/ -- CPU restatement of gordo's per-machine autoencoder anomaly path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the CPU arm that is timed beside the
GPU.  Nothing under ``gordo_b200/`` imports it; the product fails loudly when the CUDA
library is missing.

What is restated (file:line are relative to the reference checkout, equinor/gordo @ 99a4819d):

* ``factories.py``  -- gordo/machine/model/factories/utils.py:7-41 (hourglass_calc_dims),
  feedforward_autoencoder.py:15-251, lstm_autoencoder.py:15-263 (topologies only).
* ``dense.py``      -- the [3P] Keras 3.3.3 Dense forward / MSE / L1 activity regulariser /
  Adam / ``Model.fit`` batching that gordo/machine/model/models.py:243-300 delegates to.
* ``lstm.py``       -- the [3P] Keras LSTM cell + BPTT, gordo/machine/model/models.py:557-660
  (primer step, ordered batches, 10 000-window predict) and :713-793 (windowing).
* ``scaler.py``     -- [3P] sklearn MinMaxScaler (examples/config.yaml:75-82, diff.py:25).
* ``anomaly.py``    -- gordo/machine/model/anomaly/diff.py:166-458 (fit, cross_validate,
  thresholds, anomaly columns) and gordo/machine/model/utils.py:49-165 (frame layout).
* ``dataset.py``    -- the [3P] gordo-core 0.3.6 resample / interpolate / join / row-filter arithmetic
  behind ``dataset.get_data()`` (call site gordo/builder/build_model.py:208-213).

PARITY PIN STATUS
-----------------
pinned     : hourglass dims (reference golden vectors, tests/gordo/machine/model/
             test_factories_utils.py:8-24 + docstrings), windowing (test_model.py:239-311),
             the feed-forward and LSTM TOPOLOGIES (the reference's own factories executed
             against recording Keras stand-ins: tests/golden/make_topology_golden.py ->
             topology_golden.json, 17 argument sets + the factory registry), the LSTM TRAINING /
             PREDICTION PROTOCOL (which windows, in which order, primer included; the reference's
             estimator classes executed with a logging Keras model: make_protocol_golden.py),
             the builder's CV scorers and split metadata (make_metrics_golden.py),
             anomaly columns / thresholds / frame layout: checked against the REAL
             reference ``diff.py`` + ``model/utils.py`` imported here with TensorFlow stubbed
             (tests/golden/make_golden.py wrote tests/golden/*.npz), MinMaxScaler /
             TimeSeriesSplit against scikit-learn run here.
UNPINNED   : Dense / LSTM forward values, loss history, Adam trajectories: TensorFlow, Keras
             and scikeras cannot be installed in the build container (no network) and the
             reference's tests hold no golden weight / output for any Keras model.  For
             those the oracle restates the published Keras 3.3.3 algorithms; "parity
             unpinned" -- see DESIGN.md.  The hand-written backward passes are cross-checked
             against torch autograd on CPU (tests/test_oracle_grad.py), and the restated
             algorithms against a second framework's own implementations: the stacked LSTM
             forward / windowing / BPTT against ``torch.nn.LSTM``, the Dense training
             trajectory against ``torch.nn.Linear`` + ``torch.optim.Adam``
             (tests/test_oracle_torch_xcheck.py; Keras' epsilon placement is the one known
             difference).  Also pinned: DiffBasedKFCVAnomalyDetector (tests/golden/kfcv_golden.npz).
             ``dataset.py`` is UNPINNED too: gordo-core is not in the reference tree (its header says so);
             every step of it is a pandas call, which is what the GPU path is compared with.
"""
