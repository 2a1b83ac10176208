"""Extract the reference's OWN printed outputs (the executed tutorial notebook
/root/reference/examples/amrvw.ipynb, same text as README.md:51-66) into golden fixtures.
These are the only results of the reference actually *run* that exist in this image (no Julia
here).  Run in the build container only:  python tests/golden/make_golden.py
"""
import json
import os
import re

NB = "/root/reference/examples/amrvw.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_notebook.json")

num = r"[-+]?\d+\.?\d*(?:e[-+]?\d+)?"
cre = re.compile(rf"\s*({num})\s*([-+])\s*({num})im")


def parse_complex_block(txt):
    out = []
    for line in txt.splitlines()[1:]:
        m = cre.match(line)
        if m:
            re_, sg, im_ = m.groups()
            out.append([re_, ("-" if sg == "-" else "") + im_])   # keep strings: exact decimal -> float64 round trip
    return out


def main():
    nb = json.load(open(NB))
    cells = []
    for c in nb["cells"]:
        if c["cell_type"] != "code":
            continue
        src = "".join(c["source"])
        txt = ""
        for o in c.get("outputs", []):
            if "text" in o:
                txt += "".join(o["text"])
            elif "data" in o and "text/plain" in o["data"]:
                txt += "".join(o["data"]["text/plain"])
        cells.append((src, txt))

    def find(snippet, nth=0):
        hits = [t for s, t in cells if snippet in s]
        return hits[nth]

    g = {"source": "examples/amrvw.ipynb (executed outputs), README.md:51-66"}
    g["p4"] = {"coeffs": [24.0, -50.0, 35.0, -10.0, 1.0], "roots": parse_complex_block(find("A.roots(ps)", 0)),
               "cite": "examples/amrvw.ipynb cell 4 / README.md:62-66", "bitwise": True}
    g["pc"] = {"coeffs": [[0.0, 1.0], [-1.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, -1.0], [1.0, 0.0]],
               "roots": parse_complex_block(find("ps  =  [0.0 + 1.0im", 0)), "cite": "examples/amrvw.ipynb cell 8", "bitwise": True}
    g["p4_pencil"] = {"coeffs": [24.0, -50.0, 35.0, -10.0, 1.0], "roots": parse_complex_block(find("A.roots(vs, ws)", 0)),
                      "cite": "examples/amrvw.ipynb cell 70", "bitwise": True}
    g["wilkinson20"] = {"roots": parse_complex_block(find("Polynomials.poly(1.0:20.0)", 0)), "cite": "examples/amrvw.ipynb cell 73",
                        "bitwise": False, "note": "takes one exceptional (Base.rand) shift: tolerance only"}
    g["wilkinson20_pencil13"] = {"roots": parse_complex_block(find("A.roots(pencil_split(ps, 13)...)", 0)), "cite": "examples/amrvw.ipynb cell 77",
                                 "bitwise": False}
    btxt = find("B = F.RF.B", 0)
    g["p4_B_chain"] = {"rotators": re.findall(rf"Rotator\{{Float64,Float64\}}\(({num}), ({num}), (\d)\)", btxt), "cite": "examples/amrvw.ipynb cell 55"}
    g["p4_rho"] = {"value": find("rho = (en1'", 0).strip(), "cite": "examples/amrvw.ipynb cell 59"}
    json.dump(g, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
