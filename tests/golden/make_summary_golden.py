"""Extract the reference's saved `model.summary()` of the 1-stack hourglass (dev/making_hourglass.ipynb, cell 3) into
tests/golden/keras_summary_1stack.json: [[layer name, class, param count, [inbound layer names]], ...] in `model.layers`
order.  This is the only place the reference records Keras' layer ORDER (which fixes the `layer_with_weights-N` keys of its
TF checkpoints) and the auto-generated layer names.  Run in the build container only: python tests/golden/make_summary_golden.py
"""
import json
import os
import re

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    nb = json.load(open("/root/reference/dev/making_hourglass.ipynb"))
    text = "".join(nb["cells"][3]["outputs"][0]["text"])
    lines = text.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("====")) + 1
    end = next(i for i, l in enumerate(lines) if l.startswith("====") and i > start)
    rows, cur = [], None
    for l in lines[start:end]:
        if not l.strip():
            if cur:
                rows.append(cur)
            cur = None
            continue
        l = l.ljust(100)
        cols = [l[0:32], l[32:53], l[53:65], l[65:]]
        if cur is None:
            cur = ["", "", "", ""]
        for k in range(4):
            cur[k] += cols[k].strip() if k != 3 else cols[k].rstrip()[1:] if cols[k].startswith(" ") else cols[k].rstrip()
    if cur:
        rows.append(cur)
    out = []
    for name_type, _shape, params, conn in rows:
        m = re.match(r"^(\S+?)\s*\((\w+)\)$", name_type.replace(" ", ""))
        assert m, name_type
        inbound = re.findall(r"'([^'\[]+)\[0\]\[0\]'", conn.replace(" ", ""))
        out.append([m.group(1), m.group(2), int(params), inbound])
    total = int(re.search(r"Total params: ([\d,]+)", text).group(1).replace(",", ""))
    assert sum(r[2] for r in out) == total == 3659665, (sum(r[2] for r in out), total)
    json.dump(out, open(os.path.join(OUT, "keras_summary_1stack.json"), "w"), indent=0)
    print(len(out), "layers;", out[0], out[10], out[-1])


if __name__ == "__main__":
    main()
