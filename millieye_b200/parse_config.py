"""Darknet .cfg / .data readers with the reference's semantics
(module3_our_dataset/utils/parse_config.py:3-38): same block dicts, values kept as strings."""


def parse_model_config(path):
    """[net] first, then one dict per layer block. 'convolutional' blocks default to
    batch_normalize=0 (reference :14-15); comment lines must start with '#' in column 0 (:7)."""
    blocks = []
    with open(path, "r") as fh:
        for raw in fh.read().split("\n"):
            if raw == "" or raw[0] == "#":
                continue
            text = raw.strip()
            if text[:1] == "[":
                block = {"type": text[1:-1].rstrip()}
                if block["type"] == "convolutional":
                    block["batch_normalize"] = 0
                blocks.append(block)
                continue
            name, value = text.split("=")
            blocks[-1][name.rstrip()] = value.strip()
    return blocks


def parse_data_config(path):
    """key=value file with the reference's defaults (:25-27) and space-split list values (:35-36)."""
    options = {"gpus": "0,1,2,3", "num_workers": "10"}
    with open(path, "r") as fh:
        for raw in fh.readlines():
            text = raw.strip()
            if text == "" or text.startswith("#"):
                continue
            name, value = text.split("=")
            parts = value.split(" ")
            options[name.strip()] = parts if len(parts) > 1 else value
    return options
