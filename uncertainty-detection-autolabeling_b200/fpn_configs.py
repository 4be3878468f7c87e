"""Mirror of the reference's ``src/fpn_configs.py`` for the BiFPN graph (bifpn_config, :27-78): which node reads which.

A config is a plain dict ``{"weight_method": str, "nodes": [{"feat_level": int, "inputs_offsets": [int, ...]}, ...]}``;
node ids count the input features first (min_level .. max_level = 0 .. L-1), every new node appends one id.
"""


def bifpn_config(min_level, max_level, weight_method=None):
    """Top-down path P(max-1)' .. P(min)", then bottom-up path P(min+1)" .. P(max)" (fpn_configs.py:27-78)."""
    num_levels = max_level - min_level + 1
    ids = {min_level + i: [i] for i in range(num_levels)}
    nxt = num_levels
    nodes = []
    for lvl in range(max_level - 1, min_level - 1, -1):
        nodes.append({"feat_level": lvl, "inputs_offsets": [ids[lvl][-1], ids[lvl + 1][-1]]})
        ids[lvl].append(nxt)
        nxt += 1
    for lvl in range(min_level + 1, max_level + 1):
        nodes.append({"feat_level": lvl, "inputs_offsets": list(ids[lvl]) + [ids[lvl - 1][-1]]})
        ids[lvl].append(nxt)
        nxt += 1
    return {"weight_method": weight_method or "fastattn", "nodes": nodes}


def get_fpn_config(fpn_name, min_level, max_level, weight_method):
    """fpn_configs.py get_fpn_config: only the BiFPN family is on this path (``None`` / "bifpn_dyn" / "bifpn")."""
    if not fpn_name or fpn_name in ("bifpn", "bifpn_dyn"):
        return bifpn_config(min_level, max_level, weight_method)
    raise ValueError("fpn_name {} is not supported (bifpn only).".format(fpn_name))
