"""Oracle restatement of the darknet .cfg parser (reference utils/parse_config.py:3-21).

Test infrastructure only (see oracle/__init__.py).
"""


def parse_model_config(path):
    """Returns the list of block dicts; block 0 is [net].

    Rules kept from the reference: blank lines and lines starting with '#' are dropped
    (parse_config.py:7), lines are stripped (:8), '[type]' opens a block (:11-13), a
    convolutional block defaults batch_normalize to 0 (:14-15), every other line is split at
    '=' with both sides stripped (:17-19); values stay strings.
    """
    with open(path, "r") as fh:
        raw = fh.read().split("\n")
    blocks = []
    for line in raw:
        if not line or line.startswith("#"):
            continue
        line = line.strip()
        if line.startswith("["):
            blocks.append({"type": line[1:-1].rstrip()})
            if blocks[-1]["type"] == "convolutional":
                blocks[-1]["batch_normalize"] = 0
        else:
            key, value = line.split("=")
            blocks[-1][key.rstrip()] = value.strip()
    return blocks
