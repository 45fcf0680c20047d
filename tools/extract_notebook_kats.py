#!/usr/bin/env python
"""Extract the known-answer values that pin this path from the reference's EXECUTED notebook
``/root/reference/docs/tutorials/visualisation/gala.ipynb`` into tests/golden/gala_notebook_kats.json.

The reference's own tests contain no numeric orbit fixtures; these printed cell
outputs are the only numbers in the repository produced by real gala on this path:
  KAT-1  MilkyWayPotential(v2).circular_velocity([8.5,0,0] kpc)         (cell "circular_velocity")
  KAT-2  component order / disk repr
  KAT-4  time grid of a kicked orbit (timestep_size = 0.1 Myr): first/last three nodes, length
  KAT-5  first three and last three phase-space samples of that DOPRI853 orbit
Run in the build container (needs /root/reference); the JSON is committed, the GPU box never reads
the reference.
"""
import json
import os
import re

NB = "/root/reference/docs/tutorials/visualisation/gala.ipynb"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "gala_notebook_kats.json")

nb = json.load(open(NB))
text = {}
for i, c in enumerate(nb["cells"]):
    if c["cell_type"] != "code":
        continue
    src = "".join(c["source"])
    outs = []
    for o in c.get("outputs", []):
        if "text" in o:
            outs.append("".join(o["text"]))
        elif "data" in o and "text/plain" in o["data"]:
            outs.append("".join(o["data"]["text/plain"]))
    text[src] = "\n".join(outs)


def find(key):
    for src, out in text.items():
        if key in src:
            return src, out
    raise KeyError(key)


num = r"[-+]?\d+\.\d*(?:[eE][-+]?\d+)?"
src, out = find("circular_velocity(q=[8.5, 0, 0]")
vcirc = float(re.search(r"\[(" + num + r")\]", out).group(1))
src, out = find("p.galactic_potential")
order = re.search(r"<CompositePotential ([\w,]+)>", out).group(1).split(",")
_, disk = find('galactic_potential["disk"]')
src, out = find("p.orbits[0].pos, p.orbits[0].vel, p.orbits[0].t")
triples = re.findall(r"\(\s*(" + num + r")\s*,\s*(" + num + r")\s*,\s*(" + num + r")\s*\)", out)
triples = [[float(v) for v in t] for t in triples]
assert len(triples) == 12, len(triples)
pos, vel = triples[:6], triples[6:]
tline = out[out.index("<Quantity ["):]
times = [float(v) for v in re.findall(num, tline)]
assert len(times) == 6
shape_out = ""
for s, o in text.items():
    if s.rstrip().endswith("p.orbits[0]"):
        shape_out = o
n_samples = int(re.search(r"shape=\((\d+),\)", shape_out).group(1))
src_pop, _ = find("cogsworth.pop.Population(100")
dt_myr = float(re.search(r"timestep_size=([\d.]+) \* u.Myr", src_pop).group(1))

kats = {
    "source": "docs/tutorials/visualisation/gala.ipynb (executed outputs; gala + cogsworth run by the reference authors)",
    "KAT1_vcirc_kms_at_8p5kpc_v2": vcirc,
    "KAT2_component_order": order,
    "KAT2_disk_repr": disk.strip(),
    "KAT4_timestep_myr": dt_myr,
    "KAT4_n_samples": n_samples,
    "KAT4_t_first3_myr": times[:3],
    "KAT4_t_last3_myr": times[3:],
    "KAT5_pos_first3_kpc": pos[:3],
    "KAT5_pos_last3_kpc": pos[3:],
    "KAT5_vel_first3_kpc_per_myr": vel[:3],
    "KAT5_vel_last3_kpc_per_myr": vel[3:],
    "print_precision": 1e-8,
}
os.makedirs(os.path.dirname(OUT), exist_ok=True)
with open(OUT, "w") as f:
    json.dump(kats, f, indent=1)
print(json.dumps(kats, indent=1))
