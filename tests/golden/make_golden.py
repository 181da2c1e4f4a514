"""Harvest the reference's stored notebook outputs into golden fixtures.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Writes tests/golden/algebraic.json and tests/golden/predictions_example1.json.

Sources (the reference has no tests; these stored outputs are its only
result-pinning artefacts):
  examples/Algebraic-equations/algebraic_equations_graph_discovery.ipynb
  examples/Algebraic-equations/predictions-example1.ipynb
Line formats come from MAIN:750-757, 760, 772-774, 815 and 981-1002.
"""
import json
import os
import re
import sys

REF = "/root/reference/examples/Algebraic-equations"
HERE = os.path.dirname(os.path.abspath(__file__))

RE_RES = re.compile(r"^Results for (.+)$")
RE_K = re.compile(r"^Kernel \[(\w+)\] has n/\(n\+s\)=([^,]+), Z=\(([^,]+), ([^)]+)\), gamma=(\S+)$")
RE_NOANC = re.compile(r"^(.+) has no ancestors$")
RE_HAS = re.compile(r"^(.+) has ancestors with the kernel \[(\w+)\] \| \(n/\(s\+n\)=([0-9.eE+-]+) after pruning\)$")
RE_ANC = re.compile(r"^Ancestors of (.+): \[(.*)\]$")
RE_DEP = re.compile(r"^ [=+] ([0-9.]+) \* (\S+) \{kernel=(\w+), gamma=([^,]+), noise=([^}]+)\}$")


def cell_stream(cell):
    out = ""
    for o in cell.get("outputs", []):
        if o.get("output_type") == "stream":
            out += "".join(o["text"])
    return out


def parse_fit(text):
    nodes = {}
    order = []
    cur = None
    for line in text.splitlines():
        line = line.rstrip()
        m = RE_RES.match(line)
        if m:
            cur = m.group(1)
            nodes[cur] = {"stage1": {}, "ancestors": None, "kernel": None}
            order.append(cur)
            continue
        m = RE_K.match(line)
        if m and cur:
            nodes[cur]["stage1"][m.group(1)] = {
                "noise": float(m.group(2)), "noise_str": m.group(2),
                "Z_low_2dp": m.group(3), "Z_high_2dp": m.group(4), "gamma_print": m.group(5)}
            continue
        m = RE_HAS.match(line)
        if m:
            nodes[m.group(1)]["kernel"] = m.group(2)
            nodes[m.group(1)]["noise_after_2dp"] = m.group(3)
            continue
        m = RE_NOANC.match(line)
        if m and not line.startswith("Ancestors"):
            nodes[m.group(1)]["kernel"] = None
            nodes[m.group(1)]["ancestors"] = []
            continue
        m = RE_ANC.match(line)
        if m:
            anc = [s.strip().strip("'") for s in m.group(2).split(",") if s.strip()]
            nodes[m.group(1)]["ancestors"] = anc
    return {"order": order, "nodes": nodes}


def parse_deps(text):
    out = {}
    cur = None
    for line in text.splitlines():
        m = RE_DEP.match(line.rstrip())
        if m:
            out[cur]["ancestors"].append(m.group(2))
            out[cur].update(kernel=m.group(3), gamma_3sf=m.group(4), noise_3sf=m.group(5))
        elif line.strip() and not line.startswith(" "):
            cur = line.strip()
            out[cur] = {"ancestors": []}
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not present; fixtures are committed, nothing to do")
    nb = json.load(open(os.path.join(REF, "algebraic_equations_graph_discovery.ipynb")))
    fits, deps = [], []
    for c in nb["cells"]:
        if c["cell_type"] != "code":
            continue
        src = "".join(c["source"])
        txt = cell_stream(c)
        if ".fit()" in src and "Results for" in txt:
            fits.append(parse_fit(txt))
        if "show_functional_dependencies" in src:
            deps.append(parse_deps(txt))
    assert len(fits) == 4 and len(deps) == 4, (len(fits), len(deps))
    recipes = [
        "ex1: key=PRNGKey(0); key,sub=split(key); W=normal(sub,(1000,4)); X=[W[:,:2],W]",
        "ex2: next split; x1=w1; x2=x1^2+1+0.1*w2; x3=w3; X=[x1,x2,x3,W]",
        "ex3: next split; x1=w1*w2; x2=w2*sin(w4); X=[x1,x2,W]",
        "ex4: next split; x1=w1; x2=x1^3+1+0.1*w2; x3=(x1+2)^3+0.1*w3; X=[x1,x2,x3,W]",
    ]
    out = {"source": "algebraic_equations_graph_discovery.ipynb (stored outputs)",
           "examples": [{"recipe": r, "fit": f, "dependencies": d} for r, f, d in zip(recipes, fits, deps)]}
    n_vals = sum(len(nd["stage1"]) for f in fits for nd in f["nodes"].values())
    assert n_vals == 78, n_vals  # 26 nodes x 3 kernels
    json.dump(out, open(os.path.join(HERE, "algebraic.json"), "w"), indent=1)

    nb = json.load(open(os.path.join(REF, "predictions-example1.ipynb")))
    fits = []
    srcs = []
    for c in nb["cells"]:
        if c["cell_type"] != "code":
            continue
        txt = cell_stream(c)
        srcs.append("".join(c["source"]))
        if "Results for" in txt:
            fits.append(parse_fit(txt))
    assert len(fits) == 1
    out = {"source": "predictions-example1.ipynb (stored outputs)",
           "recipe": "key=PRNGKey(0); key,sub=split(key); W=normal(sub,(420,4)); X=[W[:,:2],W]; "
                     "kernels=[Linear, Quadratic, Gaussian(l=0.5)]",
           "fit": fits[0]}
    json.dump(out, open(os.path.join(HERE, "predictions_example1.json"), "w"), indent=1)
    print("wrote fixtures:", n_vals, "+", sum(len(nd["stage1"]) for nd in fits[0]["nodes"].values()), "stage-1 values")


if __name__ == "__main__":
    main()
