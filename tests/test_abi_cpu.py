"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/udal.h declares, the ctypes struct mirrors the C struct, host logic behaves like the
reference.  No compute calls - there is no GPU here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "udal.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(udal_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import udal_b200 as u
    lib = ctypes.CDLL(u._lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libudal.so does not export %s" % s
    # and the binding table covers the header one to one
    assert sorted(u._lib.SIGNATURES) == syms


def test_config_struct_matches_c_layout(tmp_path):
    import udal_b200 as u
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "udal.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(udal_config),'
                   ' offsetof(udal_config, anchors_per_loc), offsetof(udal_config, nms_score_thresh), offsetof(udal_config, inv_keep_box));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o1, o2, o3 = map(int, subprocess.check_output([str(exe)]).split())
    C = u._lib.Config
    assert ctypes.sizeof(C) == size
    assert C.anchors_per_loc.offset == o1 and C.nms_score_thresh.offset == o2 and C.inv_keep_box.offset == o3


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "udal.h"\nint main(void){return UDAL_ABI_VERSION - 1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_gpu_fails_loudly_not_silently():
    import udal_b200 as u
    if u._lib.device_count() > 0:
        pytest.skip("GPU present")
    p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=64, num_classes=3,
                                              enable_softmax=True)
    cls = [np.zeros((1, 8 >> i or 1, 8 >> i or 1, 27), np.float32) for i in range(5)]
    with pytest.raises((RuntimeError, ValueError)):
        u.postprocess.postprocess_global(p, cls, cls)


def test_missing_library_is_an_import_error(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import importlib\n"
            "m = importlib.import_module('uncertainty-detection-autolabeling_b200._lib')\n"
            "m.LIB_PATH = %r; m._lib = None\n"
            "try:\n    m.load()\nexcept ImportError as e:\n    print('IMPORT_ERROR'); sys.exit(0)\nsys.exit(1)\n"
            % (ROOT, str(tmp_path / "nope.so")))
    out = subprocess.check_output([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=ROOT))
    assert b"IMPORT_ERROR" in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uncertainty-detection-autolabeling_b200")
    bad = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.+oracle|#\s*include\s*[\"<].*oracle)|liboracle|oracle/_build", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f


def test_host_helpers_match_reference_semantics():
    import udal_b200 as u
    assert u.utils.parse_image_size("1024x512") == (512, 1024)
    assert u.utils.parse_image_size(512) == (512, 512)
    assert u.utils.parse_image_size((3, 4)) == (3, 4)
    with pytest.raises(ValueError):
        u.utils.parse_image_size(1.5)
    fs = u.utils.get_feat_sizes((720, 1280), 7)
    assert [(f["height"], f["width"]) for f in fs[3:]] == [(90, 160), (45, 80), (23, 40), (12, 20), (6, 10)]
    assert u.postprocess.to_list({"b": 2, "a": 1}) == [1, 2]
    assert u.postprocess.to_list((1, 2)) == [1, 2]
    assert u.postprocess.to_list(3) is None
    with pytest.raises(ValueError, match="invalid nms method"):
        u.engine.nms_thresholds(dict(method="bogus", iou_thresh=None, score_thresh=None, sigma=None))
    assert u.engine.nms_thresholds(dict(method="gaussian", iou_thresh=None, score_thresh=0.0, sigma=None)) == (1, 0.25, 0.5, 0.001)
    assert u.engine.nms_thresholds(dict(method="hard", iou_thresh=None, score_thresh=None, sigma=None))[1:] == (0.0, 0.5, float("-inf"))


def test_anchor_table_matches_golden():
    import udal_b200 as u
    from tests.helpers import load_golden
    g = load_golden("anchors")
    for tag, size in (("512", 512), ("384x1280", (384, 1280)), ("720x1280", (720, 1280)), ("768", 768),
                      ("str1024x512", "1024x512"), ("64x96", (64, 96))):
        a = u.anchors.Anchors(3, 7, 3, [1.0, 2.0, 0.5], 4.0, size)
        assert a.boxes.shape[0] == int(g[tag + "_n"])
        np.testing.assert_array_equal(a.boxes[g[tag + "_rows"]], g[tag + "_vals"])
        np.testing.assert_array_equal(a.boxes.astype(np.float64).sum(0), g[tag + "_sum64"])
        assert a.get_anchors_per_location() == 9
        # the table's raw pointer is uploaded to the device: it must be C-contiguous [N,4] fp32 at EVERY geometry (round 1
        # uploaded a Fortran-ordered table for every level above ~100 px: NumPy's concatenate / astype kept the order of
        # the transposed views)
        t = u.engine.anchor_table(3, 7, 3, [1.0, 2.0, 0.5], 4.0, size)
        assert t.flags["C_CONTIGUOUS"] and t.dtype == np.float32 and t.strides == (16, 4)
        np.testing.assert_array_equal(np.frombuffer(t.tobytes(), np.float32).reshape(-1, 4), a.boxes)
    np.testing.assert_array_equal(
        u.anchors.Anchors(3, 5, 2, [1.0, [1.4, 0.7]], [4.0, 3.0, 5.0], 128).boxes, g["custom_128"])


def test_philox_reference_vector():
    # Random123 known answer: philox4x32-10, counter = key = 0
    import udal_b200 as u
    m = u.heads.philox_keep_masks((1, 2, 1, 1, 1, 4), 0.0, 0.0, 0)
    assert m.shape == (1, 2, 1, 1, 1, 4) and m.dtype == np.uint8 and m.all()
    # statistical sanity of the uniform stream
    k = u.heads.philox_keep_masks((8, 2, 5, 3, 4, 64), 0.05, 0.3, 1234)
    assert abs(k[:, 0].mean() - 0.95) < 0.01 and abs(k[:, 1].mean() - 0.7) < 0.01


def test_autolabel_params_struct_matches_c_layout(tmp_path):
    import udal_b200 as u
    src = tmp_path / "sz2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "udal.h"\nint main(){printf("%zu %zu %zu %zu\\n", '
                   'sizeof(udal_autolabel_params), offsetof(udal_autolabel_params, table_off), '
                   'offsetof(udal_autolabel_params, class_temp), offsetof(udal_autolabel_params, strict_reference));return 0;}\n')
    exe = tmp_path / "sz2"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o1, o2, o3 = map(int, subprocess.check_output([str(exe)]).split())
    P = u._lib.AutolabelParams
    assert ctypes.sizeof(P) == size
    assert P.table_off.offset == o1 and P.class_temp.offset == o2 and P.strict_reference.offset == o3


def test_synthetic_generators_match_the_oracle_copy():
    """bench.py / tools take their synthetic weights, features and masks from the package (the product never imports
    oracle/); the oracle keeps its own copy of the generators - both must give the same arrays for the same seeds."""
    import importlib.util
    import os
    import numpy as np
    from oracle import heads_ref
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "uncertainty-detection-autolabeling_b200",
                        "synthetic.py")
    spec = importlib.util.spec_from_file_location("udal_synthetic_standalone", path)  # no libudal needed
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    for kw in (dict(seed=2024), dict(seed=9, randomize_bn=True)):
        a = syn.init_head_weights(64, 3, 5, 9, 8, True, **kw)
        b = heads_ref.init_head_weights(64, 3, 5, 9, 8, True, **kw)
        for head in ("class", "box"):
            for key in ("dw", "pw", "b"):
                for x, y in zip(a[head][key], b[head][key]):
                    np.testing.assert_array_equal(x, y)
            for key in ("dwp", "pwp", "bp"):
                np.testing.assert_array_equal(a[head][key], b[head][key])
            for ra, rb in zip(a[head]["bn"], b[head]["bn"]):
                for la, lb in zip(ra, rb):
                    for key in la:
                        np.testing.assert_array_equal(la[key], lb[key])
    np.testing.assert_array_equal(syn.make_masks(4, 5, 3, 2, 64, 0.05, 0.1, seed=5), heads_ref.make_masks(4, 5, 3, 2, 64, 0.05, 0.1, seed=5))
    for x, y in zip(syn.make_features([(4, 6), (2, 3)], 2, 64, seed=11), heads_ref.make_features([(4, 6), (2, 3)], 2, 64, seed=11)):
        np.testing.assert_array_equal(x, y)


def test_decode_precision_is_validated_on_the_host():
    """params["decode_precision"]: "fp64" (default) | "fp32"; anything else is a ValueError before a context is created,
    and the key takes part in the engine cache key (a configuration never inherits the other arithmetic's context)."""
    import udal_b200 as u
    p = u.hparams_config.get_detection_config("efficientdet-d0", image_size=64, num_classes=7, decode_precision="fp16")
    with pytest.raises(ValueError, match="decode_precision"):
        u.engine.Engine(p)
    a = u.hparams_config.get_detection_config("efficientdet-d0", image_size=64, num_classes=7)
    b = dict(a, decode_precision="fp32")
    assert u.engine._key(a, 0, "fp32") != u.engine._key(b, 0, "fp32")
    # without the key the arithmetic follows strict_reference (default True -> the reference's float64 decode)
    assert u.engine.decode_precision(a) == "fp64" and u.engine.decode_precision(dict(a, strict_reference=False)) == "fp32"
    assert u.engine.decode_precision(dict(a, strict_reference=False, decode_precision="fp64")) == "fp64"
    assert _lib_config_field("decode_precision")


def _lib_config_field(name):
    import udal_b200 as u
    return name in [f[0] for f in u._lib.Config._fields_]


def test_product_and_tools_do_not_import_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU baseline legs may use it."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for sub in ("uncertainty-detection-autolabeling_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith(".py") and pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.join(sub, f))
    assert not offenders, offenders
    # bench.py: the oracle appears only inside the CPU arms (cpu_port_step / run_reference / the cpu_baseline leg)
    src = open(os.path.join(root, "bench.py")).read()
    gpu_arm = src[src.index("def run_gpu("):]
    lines = [l for l in gpu_arm.splitlines() if pat.search(l)]
    assert all("oracle_build" in l for l in lines), lines   # building the checker for the cpu_baseline leg is not using it


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: ONE JSON line on stdout with the GPU arm's metric / unit / config keys, impl = reference,
    a cpu_baseline describing the run and an e2e block without copies (runs the oracle port on 2 images per step)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mc_dropout_images_per_sec_effdet_d0_T10" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("EfficientDet-D0 1280x384") and d["config"]["T"] == 10
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
