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
