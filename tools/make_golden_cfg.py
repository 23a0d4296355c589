"""Merge the reference's two-deck test configs (defaults ⊕ inputs, runner.py:70-72 semantics:
flatten -> update -> unflatten) into single JSON decks under tests/golden/ so that tests can run
where /root/reference does not exist.  Mirrors what tests/test_forward/test_1d.py:33-52 does."""
import json, os, yaml

REF = "/root/reference/tests/configs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def flatten(d, pre=()):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict) and v:
            out.update(flatten(v, pre + (k,)))
        else:
            out[pre + (k,)] = v
    return out


def unflatten(f):
    out = {}
    for ks, v in f.items():
        d = out
        for k in ks[:-1]:
            d = d.setdefault(k, {})
        d[ks[-1]] = v
    return out


def merged(defaults, inputs):
    with open(os.path.join(REF, defaults)) as fi:
        d = flatten(yaml.safe_load(fi))
    with open(os.path.join(REF, inputs)) as fi:
        d.update(flatten(yaml.safe_load(fi)))
    cfg = unflatten(d)
    fr = cfg["data"]["fit_rng"]
    cfg["other"]["lamrangE"] = [fr["forward_epw_start"], fr["forward_epw_end"]]
    cfg["other"]["lamrangI"] = [fr["forward_iaw_start"], fr["forward_iaw_end"]]
    cfg["other"]["npts"] = int(cfg["other"]["CCDsize"][1] * cfg["other"]["points_per_pixel"])
    return cfg


for name, (d, i) in {
    "cfg_1d": ("1d-defaults.yaml", "1d-inputs.yaml"),
    "cfg_epw": ("epw_defaults.yaml", "epw_inputs.yaml"),
    "cfg_arts1v": ("arts1v_test_defaults.yaml", "arts1v_test_inputs.yaml"),
    "cfg_arts2v": ("arts2v_test_defaults.yaml", "arts2d_test_inputs.yaml"),
}.items():
    with open(os.path.join(OUT, name + ".json"), "w") as fo:
        json.dump(merged(d, i), fo, indent=1, sort_keys=True)
    print(name)
