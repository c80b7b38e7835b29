"""ORACLE tooling (test infrastructure): extract the reference's own `print(model)` dump into a committed fixture.

The only artefact in the reference repository that shows what `timm.create_model("mobilenetv4_conv_small",
features_only=True)` builds is the training log saved in YoloLite_custom_training.ipynb (code cell 8's stdout,
notebook file lines ~391-1006): the full module tree of the edge_s model (YOLOLiteMS_CPU, mobilenetv4_conv_small
backbone, fpn 192, depth 2, head_depth 2, P6 enabled, 13 classes).  This script copies that block VERBATIM into
tests/golden/notebook_model_dump.txt so that tests/test_notebook_pin.py can compare the oracle's and the packer's
layer tables with it on machines where /root/reference does not exist.

    python oracle/extract_notebook_dump.py [/root/reference]
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def extract(ref_root: str) -> str:
    with open(os.path.join(ref_root, "YoloLite_custom_training.ipynb")) as f:
        nb = json.load(f)
    for ci, cell in enumerate(nb["cells"]):
        for out in cell.get("outputs", []):
            lines = "".join(out.get("text", [])).split("\n")
            starts = [i for i, l in enumerate(lines) if l.startswith("YOLOLiteMS_CPU(")]
            if not starts:
                continue
            i0 = starts[0]
            i1 = next(i for i in range(i0, len(lines)) if lines[i] == ")")
            head = (f"# verbatim copy of the `print(model)` block in YoloLite_custom_training.ipynb (cell {ci}, stdout lines "
                    f"{i0}-{i1}); extracted by oracle/extract_notebook_dump.py\n")
            return head + "\n".join(lines[i0:i1 + 1]) + "\n"
    raise RuntimeError("no YOLOLiteMS_CPU( block found in the notebook outputs")


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dst = os.path.join(REPO, "tests", "golden", "notebook_model_dump.txt")
    txt = extract(root)
    with open(dst, "w") as f:
        f.write(txt)
    print(f"wrote {dst}: {txt.count(chr(10))} lines")
