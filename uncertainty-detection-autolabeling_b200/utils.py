"""Host helpers mirrored from the reference's ``src/utils.py`` (hot-path subset).

utils.py:516-540 parse_image_size, :543-559 get_feat_sizes, :42-59 activation_fn ('swish').
"""


def parse_image_size(image_size):
    """int -> (s, s); 'WxH' string -> (H, W); (H, W) tuple unchanged (utils.py:516-540)."""
    if isinstance(image_size, int):
        return (image_size, image_size)
    if isinstance(image_size, str):
        width, height = image_size.lower().split("x")
        return (int(height), int(width))
    if isinstance(image_size, tuple):
        return image_size
    raise ValueError(
        "image_size must be an int, WxH string, or (height, width)tuple. Was %r" % (image_size,))


def get_feat_sizes(image_size, max_level):
    """[{'height': h, 'width': w}] for levels 0..max_level (utils.py:543-559)."""
    h, w = parse_image_size(image_size)
    sizes = [{"height": h, "width": w}]
    for _ in range(1, max_level + 1):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        sizes.append({"height": h, "width": w})
    return sizes


def mix_seed(seed, stream):
    """64-bit seed of sub-stream ``stream`` (a shard, a batch of a map) of ``seed``: splitmix64 of the pair, so that
    callers stepping ``seed`` by one never collide with the sub-streams of a neighbouring call (seed + i would)."""
    m = (1 << 64) - 1
    z = (int(seed) * 0x9E3779B97F4A7C15 + (int(stream) + 1) * 0xD1B54A32D192ED03) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)
