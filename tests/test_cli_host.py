"""Host driver (nfx.cli): the reference's argv contract, path validation, GeoJSON model and writers.
CPU part: parsing/validation/IO. GPU part: end-to-end run compared with the oracle pipeline."""
import json
import os

import numpy as np
import pytest

import nfx
import nfx_oracle as o
from nfx import cli, synth


def _write_geojson(path, rings, with_bbox=True):
    feats = []
    for r in rings:
        ft = {"type": "Feature", "geometry": {"type": "Polygon", "coordinates": [[[float(x), float(y)] for x, y in r]]}}
        if with_bbox:
            ft["bbox"] = [float(r[:, 0].min()), float(r[:, 1].min()), float(r[:, 0].max()), float(r[:, 1].max())]
        feats.append(ft)
    json.dump({"type": "FeatureCollection", "features": feats}, open(path, "w"))


def test_argv_contract_and_defaults():
    a = cli.build_parser().parse_args(["g.geojson", "s.png", "out.csv", "color", "glcm"])
    assert (a.patch_size, a.batch_size, a.overwrite, a.verbose, a.gpus, a.thread_count) == (64, 100, False, False, None, None)
    assert a.feature_sets == ["color", "glcm"]
    a = cli.build_parser().parse_args(["-o", "-p", "32", "-b", "50", "-g", "0", "1", "-v", "--", "g", "s", "o.pqt", "all"])
    assert (a.patch_size, a.batch_size, a.overwrite, a.verbose, a.gpus) == (32, 50, True, True, [0, 1])


def test_validate_paths_matches_reference_messages(tmp_path, caplog):
    g, s = tmp_path / "g.geojson", tmp_path / "s.png"
    g.write_text("{}")
    s.write_bytes(b"x")

    def run(args):
        with pytest.raises(SystemExit) as e:
            cli.validate_paths(cli.build_parser().parse_args(args))
        assert e.value.code == 1
        return caplog.text

    assert "Geometry file does not exist" in run([str(tmp_path / "nope"), str(s), "o.csv", "all"])
    assert "Slide file does not exist" in run([str(g), str(tmp_path / "nope.png"), "o.csv", "all"])
    assert "Output file must have an extension" in run([str(g), str(s), str(tmp_path / "out"), "all"])
    assert "Unsupported output format" in run([str(g), str(s), str(tmp_path / "out.xlsx"), "all"])
    out = tmp_path / "out.csv"
    out.write_text("x")
    assert "Output file already exists" in run([str(g), str(s), str(out), "all"])
    assert cli.validate_paths(cli.build_parser().parse_args(["-o", str(g), str(s), str(out), "all"])) == "csv"


def test_geojson_model_f32_ring0_bbox_required(tmp_path):
    rings = [np.array([[0.1, 0.2], [10.123456789, 0.2], [5, 7.5], [0.1, 0.2]], np.float64)]
    p = tmp_path / "a.geojson"
    _write_geojson(p, rings)
    xy, off = cli.load_geometry(p)
    assert xy.dtype == np.float32 and off.tolist() == [0, 4]
    assert np.array_equal(xy, rings[0].astype(np.float32))          # parsed AS f32, closing duplicate kept
    _write_geojson(p, rings, with_bbox=False)
    with pytest.raises(nfx.NfxError, match="missing field `bbox`"):
        cli.load_geometry(p)                                        # src/geojson.rs:18: bbox is not Option


@pytest.mark.parametrize("ext", ["csv", "parquet", "pqt", "json", "ipc", "feather"])
def test_writers_roundtrip(tmp_path, ext):
    import pyarrow.csv as pacsv
    import pyarrow.feather as feather
    import pyarrow.parquet as pq
    keys = ["1024,33.5", "7,8"]
    feats = np.array([[1.5, np.nan], [0.25, 3.0]], np.float32)
    path = str(tmp_path / f"o.{ext}")
    cli.write_output(path, ext, keys, feats, ["mean_r", "std_r"])
    if ext == "csv":
        t = pacsv.read_csv(path)
    elif ext in ("parquet", "pqt"):
        t = pq.read_table(path)
    elif ext == "json":
        rows = [json.loads(ln) for ln in open(path)]
        assert [r["centroid"] for r in rows] == keys and rows[1]["std_r"] == 3.0 and rows[0]["std_r"] is None
        return
    else:
        t = feather.read_table(path)
    assert t.column_names == ["centroid", "mean_r", "std_r"]
    assert t.column("centroid").to_pylist() == keys
    assert t.column("mean_r").to_pylist() == [1.5, 0.25]


@pytest.mark.gpu
def test_cli_end_to_end_matches_reference_pipeline(tmp_path, libnfx):
    from PIL import Image
    import pyarrow.parquet as pq
    tile = synth.synth_tile(384, 384, 21)
    xy, off = synth.synth_polygons(130, 384, 384, 21, border_frac=0.05)
    rings = synth.rings_of(xy, off)
    Image.fromarray(tile).save(tmp_path / "slide.png")
    _write_geojson(tmp_path / "cells.geojson", rings)
    out = tmp_path / "features.parquet"
    rc = cli.main([str(tmp_path / "cells.geojson"), str(tmp_path / "slide.png"), str(out), "geometry", "Color", "-b", "50"])
    assert rc == 0
    t = pq.read_table(out)
    # the geojson text round-trips doubles; the reference parses them to f32 exactly like load_geometry
    rings32 = [np.asarray(r, np.float64).astype(np.float32) for r in rings]
    keys, cents, want, names = o.extract(rings32, tile, ["geometry", "color"], 64, 50)
    assert t.column_names == ["centroid"] + names
    assert t.column("centroid").to_pylist() == keys
    got = np.stack([np.asarray(t.column(c).to_pylist(), dtype=np.float64) for c in ("area", "perimeter", "mean_r", "std_b", "mean_v")], 1)
    sel = [names.index(c) for c in ("area", "perimeter", "mean_r", "std_b", "mean_v")]
    assert np.allclose(got, want[:, sel], rtol=1e-4, atol=1e-6, equal_nan=True)
    # duplicate sets fail like DataFrame::new (main.rs:89); unknown names like FromStr (args.rs:29)
    with pytest.raises(SystemExit):
        cli.main(["-o", str(tmp_path / "cells.geojson"), str(tmp_path / "slide.png"), str(out), "glcm", "texture"])
