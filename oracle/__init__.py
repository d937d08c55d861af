"""
oracle/ -- CPU restatement of the DLWP forecast-rollout hot path.  TEST INFRASTRUCTURE ONLY.

This package is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``dlwp_b200/`` imports it,
and the product path fails loudly if its CUDA library is missing instead of falling back to this code.

What it restates (all citations relative to the reference checkout, jweyn/DLWP @ 3f32bfab):

* ``ops.py``      PeriodicPadding2D (DLWP/custom.py:191-214), Keras ZeroPadding2D / Conv2D / MaxPooling2D /
                  UpSampling2D semantics (Keras 2.2 API; not vendored in the reference, see "pinning" below),
                  slice_layer (DLWP/custom.py:675-692), row_conv2d (DLWP/custom.py:840-896).
* ``layers.py``   the (name, args, kwargs) layer-tuple interpreter of DLWPNeuralNet.build_model
                  (DLWP/model/models.py:63-112) and the functional ``skip_model`` / ``basic_model`` graphs of
                  examples/train_functional.py:222-285.
* ``rollout.py``  DLWPNeuralNet.predict_timeseries (DLWP/model/models.py:247-301) and
                  DLWPFunctional.predict_timeseries (DLWP/model/models.py:414-452).
* ``torch_cpu.py`` the same forward in torch-CPU fp32 (multi-threaded oneDNN): the "reference-precision" comparator and
                  the CPU timing stand-in for the reference's Keras-CPU path (Keras/TensorFlow are not installable
                  offline).

Pinning status
--------------
The reference ships no tests, golden vectors or fixtures for this path, and its Conv2D arithmetic lives in un-vendored,
un-pinned Keras/TensorFlow.  The oracle is therefore pinned against outputs of the reference ITSELF, generated in the
build container by ``tests/golden/make_golden.py`` and committed under ``tests/golden/``:

* PeriodicPadding2D: the reference's own ``call`` (custom.py:191-214) executed on numpy arrays with a stub backend;
* rollout loops: the reference's own ``DLWPNeuralNet`` / ``DLWPFunctional`` classes (models.py) executed in place
  with a duck-typed ``.model``;
* end-to-end (padding + conv + tanh + loop): the reference's own torch twin ``DLWPTorchNN`` (models_torch.py:77-376)
  running torch.nn CircularPad2d/ZeroPad2d/Conv2d on CPU.

Keras ``Conv2D`` itself cannot be executed here; its cross-correlation/"valid"/dilation semantics are restated from the
Keras 2.2 API and cross-checked against ``torch.nn.functional.conv2d`` (an independent implementation).  For that one
operator parity is "pinned by proxy" -- DESIGN.md says so.
"""
