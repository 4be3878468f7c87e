"""Generates tests/golden/bifpn_graph.json: the node graph of the reference's own ``src/fpn_configs.py`` (bifpn_config),
imported unmodified with the NumPy-backed ``tensorflow`` stand-in (its ``hparams_config`` import needs a ``tensorflow``
module to exist).  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_bifpn.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
sys.path.insert(0, "/root/reference/src")
import fpn_configs as ref  # noqa: E402

out = {}
for lo, hi in [(3, 7), (2, 7), (3, 5), (3, 4), (4, 8)]:
    for wm in (None, "sum"):
        c = ref.bifpn_config(lo, hi, wm)
        out["%d_%d_%s" % (lo, hi, wm)] = {
            "weight_method": c.weight_method,
            "nodes": [{"feat_level": int(n["feat_level"]), "inputs_offsets": [int(v) for v in n["inputs_offsets"]]} for n in c.nodes]}
path = os.path.join(HERE, "bifpn_graph.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path, len(out), "graphs")
