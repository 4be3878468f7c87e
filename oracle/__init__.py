"""CPU oracle for the uncertainty-sampling / post-processing hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``uncertainty-detection-autolabeling_b200/``) never imports, links or executes anything
from here and fails loudly when its CUDA library is missing.

What it is: a NumPy (+ torch-CPU conv for the head towers, + one small C file for the
sequential TF ``NonMaxSuppressionV5`` kernel) restatement of the reference algorithm for the
path  BiFPN features -> T x class/box(+sigma) heads -> decode with exact moment propagation
-> MC mean/std -> top-k / max-reduce -> NMS -> detections.  Every function cites the
reference file:line it follows (paths relative to the reference's ``src/``).

Parity pin status (see DESIGN.md "Oracle"):

* ``nms_np_ref``           PINNED  - checked against the reference's own ``src/nms_np.py``
                                     (NumPy only, importable in the build container); golden
                                     vectors in ``tests/golden/nms_np_*.npz``.
* ``ref_np`` (anchors, decode_uncert, get_mcuncert, merge/top-k/pre_nms/postprocess_* control
  flow)                    PINNED to the reference's *own source text*: the golden vectors in
                           ``tests/golden/`` were produced by executing the unmodified
                           ``src/postprocess.py``, ``src/anchors.py``, ``src/utils_box.py``,
                           ``src/utils_extra.py`` with a NumPy-backed stand-in for the
                           ``tensorflow`` module (``tests/golden/tf_numpy_shim.py``,
                           generator ``tests/golden/make_golden.py``).  TensorFlow itself is not
                           installable here, so the *primitive ops* (exp, top_k tie order,
                           reduce_std, gather_nd ...) are the shim's restatement.
* TF ``NonMaxSuppressionV5`` (``nms_v5.c``), ``SeparableConv2D`` / ``BatchNormalization`` /
  ``SpatialDropout2D`` (``heads_ref``)
                           PARITY UNPINNED - third-party kernels (tensorflow==2.10.0, not
                           vendored, not installable offline).  Restated from the published
                           algorithm; the reference holds no golden vectors for them.
"""
