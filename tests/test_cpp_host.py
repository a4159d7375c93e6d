"""nfx-cli, the C++ host driver (nuclei-feature-extraction_b200/host/): the reference's argv contract and
messages (src/args.rs:76-183), GeoJSON model (src/geojson.rs:8-24), PNG input, the FeatureSet trait
objects (--via-trait) and the CSV / JSON writers, against the oracle pipeline."""
import csv
import json
import os
import subprocess

import numpy as np
import pytest

import nfx_oracle as o
from nfx import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nuclei-feature-extraction_b200", "nfx-cli")


def _run(args, **kw):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=600, **kw)


def _write_geojson(path, rings, with_bbox=True):
    feats = []
    for r in rings:
        ft = {"type": "Feature", "properties": {"classification": {"name": "nucleus"}},
              "geometry": {"type": "Polygon", "coordinates": [[[float(x), float(y)] for x, y in r]]}}
        if with_bbox:
            ft["bbox"] = [float(r[:, 0].min()), float(r[:, 1].min()), float(r[:, 0].max()), float(r[:, 1].max())]
        feats.append(ft)
    json.dump({"type": "FeatureCollection", "features": feats}, open(path, "w"))


@pytest.fixture(scope="module")
def inputs(tmp_path_factory, libnfx):
    from PIL import Image
    d = tmp_path_factory.mktemp("cpp")
    tile = synth.synth_tile(320, 352, 31)
    xy, off = synth.synth_polygons(110, 320, 352, 31, border_frac=0.06)
    rings = [np.asarray(r, np.float64).astype(np.float32) for r in synth.rings_of(xy, off)]
    Image.fromarray(tile).save(d / "slide.png")
    _write_geojson(d / "cells.geojson", rings)
    return dict(dir=d, tile=tile, rings=rings)


def test_usage_and_reference_error_messages(inputs):
    d = inputs["dir"]
    r = _run([])
    assert r.returncode == 1 and "required arguments" in r.stderr
    r = _run([d / "nope.geojson", d / "slide.png", d / "o.csv", "all"])
    assert r.returncode == 1 and "Geometry file does not exist" in r.stderr
    r = _run([d / "cells.geojson", d / "nope.png", d / "o.csv", "all"])
    assert r.returncode == 1 and "Slide file does not exist" in r.stderr
    r = _run([d / "cells.geojson", d / "slide.png", d / "out", "all"])
    assert r.returncode == 1 and "Output file must have an extension" in r.stderr
    r = _run([d / "cells.geojson", d / "slide.png", d / "o.xlsx", "all"])
    assert r.returncode == 1 and "Unsupported output format" in r.stderr
    (d / "exists.csv").write_text("x")
    r = _run([d / "cells.geojson", d / "slide.png", d / "exists.csv", "all"])
    assert r.returncode == 1 and "Output file already exists" in r.stderr and "--overwrite" in r.stderr
    r = _run([d / "cells.geojson", d / "slide.png", d / "o.csv", "colour"])
    assert r.returncode == 1 and "colour is not a valid feature set" in r.stderr          # args.rs:29
    r = _run([d / "cells.geojson", d / "slide.png", d / "o.csv", "glcm", "TEXTURE"])
    assert r.returncode == 1 and "duplicate feature set" in r.stderr                      # main.rs:89
    r = _run([d / "cells.geojson", d / "slide.png", d / "o.csv"])
    assert r.returncode == 1 and "no feature set given" in r.stderr                       # main.rs:76


def test_geojson_requires_bbox_and_png_decodes(inputs):
    import torch
    d = inputs["dir"]
    _write_geojson(d / "nobbox.geojson", inputs["rings"][:3], with_bbox=False)
    r = _run([d / "nobbox.geojson", d / "slide.png", d / "o2.csv", "color"])
    assert r.returncode == 1 and "missing field `bbox`" in r.stderr                       # geojson.rs:18
    if not torch.cuda.is_available():
        # valid inputs get through parsing, the geojson and the PNG decoder, then stop at the GPU: no CPU fallback
        r = _run([d / "cells.geojson", d / "slide.png", d / "o3.csv", "color"])
        assert r.returncode == 1 and "Extracting features" in r.stderr
        assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr


def _read_csv(path):
    rows = list(csv.reader(open(path)))
    names = rows[0]
    keys = [r[0] for r in rows[1:]]
    vals = np.array([[float(x) if x != "" else np.nan for x in r[1:]] for r in rows[1:]], dtype=np.float64)
    return names, keys, vals


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--via-trait"], ["-g", "0", "0"]])
def test_cli_matches_reference_pipeline(inputs, extra):
    d = inputs["dir"]
    out = d / f"features{len(extra)}.csv"
    sets = ["geometry", "color", "glcm"]
    r = _run(["-o", "-b", "40", *extra, "--", d / "cells.geojson", d / "slide.png", out, *sets])
    assert r.returncode == 0, r.stderr
    names, keys, got = _read_csv(out)
    wkeys, wc, want, wnames = o.extract(inputs["rings"], inputs["tile"], sets, 64, 40)
    assert names == ["centroid"] + wnames and keys == wkeys
    sel = [wnames.index(c) for c in ("area", "major_axis", "perimeter", "convex_hull_area", "mean_r", "std_g", "mean_s", "mean_v",
                                     "mean_eosin", "contrast_0_1_32", "entropy_1_1_64", "sum_average_1_-1_254")]
    assert np.allclose(got[:, sel], want[:, sel], rtol=1e-4, atol=1e-6, equal_nan=True)
    if extra != ["--via-trait"]:      # mean_h: same chunks as the reference (the trait path keeps them too)
        j = wnames.index("mean_h")
        dlt = np.abs(got[:, j] - want[:, j])
        assert np.nanmax(np.minimum(dlt, 360 - dlt)) < 0.05


@pytest.mark.gpu
def test_cli_json_writer(inputs):
    d = inputs["dir"]
    out = d / "f.json"
    r = _run(["-o", d / "cells.geojson", d / "slide.png", out, "color"])
    assert r.returncode == 0, r.stderr
    rows = [json.loads(ln) for ln in open(out)]
    assert len(rows) == len(inputs["rings"]) and list(rows[0])[:3] == ["centroid", "mean_r", "mean_g"]


@pytest.mark.gpu
def test_cli_csv_device_formatter_equals_host_writer(inputs):
    """The CSV cells formatted on the GPU (nfx_csv_format) are byte-identical to the host writer's."""
    d = inputs["dir"]
    a, b = d / "dev.csv", d / "host.csv"
    for out, extra in ((a, []), (b, ["--host-csv"])):
        r = _run(["-o", *extra, "--", d / "cells.geojson", d / "slide.png", out, "all"])
        assert r.returncode == 0, r.stderr
    ta, tb = open(a, "rb").read(), open(b, "rb").read()
    assert ta == tb and ta.count(b"\n") == len(inputs["rings"]) + 1
    assert ta.split(b"\n")[1].startswith(b'"')
