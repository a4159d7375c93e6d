"""Host driver with the reference's argv contract (readme.md:10, src/args.rs:76-113):

    python -m nfx.cli [options] <input-geojson> <input-slide> <output-file> <feature-set>...

The reference's host stays Rust; rustc is not in this image, so this Python mirror plays the role
of `main()` (src/main.rs:110-190) around the C ABI for tests and demos: GeoJSON -> CSR polygons
(src/geojson.rs:8-24, coordinates parsed as f32), image load (src/main.rs:20-35), per-GPU contiguous
ranges (nfx_partition), hstack of the sets behind the `centroid` key column (src/main.rs:76-89) and
the writers by extension (src/main.rs:160-189). Feature computation itself only happens in
libnfx.so -- there is no CPU path, so `--gpus` defaults to GPU 0 instead of "use the cpu".
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
import threading

import numpy as np

log = logging.getLogger("nfx")
OUTPUT_EXT = ("csv", "parquet", "pqt", "json", "ipc", "feather")        # src/args.rs:156-157


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="nuclei-feature-extraction", description=__doc__.split("\n\n")[0])
    ap.add_argument("geometry", help="Input geometry file (.geojson)")                       # args.rs:78-79
    ap.add_argument("slide", help="Input slide file (.png/.jpg/.jpeg; .svs needs OpenSlide)")  # args.rs:80-81
    ap.add_argument("output", help="Output file: csv, parquet|pqt, json, ipc|feather")         # args.rs:82-84
    ap.add_argument("feature_sets", nargs="*", help="geometry color glcm glrlm gabor texture all")
    ap.add_argument("-o", "--overwrite", action="store_true")                                 # args.rs:88-90
    ap.add_argument("-p", "--patch-size", type=int, default=64)                               # args.rs:93-94
    ap.add_argument("-t", "--thread-count", type=int, default=None)                           # args.rs:98-99
    ap.add_argument("-g", "--gpus", type=int, nargs="*", default=None)                        # args.rs:103-104
    ap.add_argument("-b", "--batch-size", type=int, default=100)                              # args.rs:107-108
    ap.add_argument("-v", "--verbose", action="store_true")                                   # args.rs:111-112
    return ap


def die(msg: str):
    """error!(..); exit(1) (src/args.rs:137-183)."""
    log.error(msg)
    raise SystemExit(1)


def validate_paths(a):
    """Args::validate_paths (src/args.rs:137-167), same messages."""
    if not os.path.exists(a.geometry):
        die(f"Geometry file does not exist : {a.geometry!r}")
    if not os.path.exists(a.slide):
        die(f"Slide file does not exist : {a.slide!r}")
    if os.path.exists(a.output) and not a.overwrite:
        die(f"Output file already exists : {a.output!r}\nUse --overwrite to overwrite it")
    ext = os.path.splitext(a.output)[1].lstrip(".")
    if ext == "":
        die("Output file must have an extension")
    if ext not in OUTPUT_EXT:
        die("Unsupported output format. Please use one of the following : csv, parquet, json, ipc, feather")
    return ext


def load_geometry(path, threads: int = 0):
    """load_geometry (src/main.rs:37-42) + the serde model of src/geojson.rs:8-24: every feature
    needs `bbox` and `geometry.{type,coordinates}`; ring 0 is taken as stored (closing duplicate kept)
    and parsed to f32 with serde_json's number rule by the library's multi-threaded packer
    (nfx_geojson_parse). Returns CSR (poly_xy f32 [sum V, 2], poly_off int64 [n+1])."""
    import nfx
    text = np.fromfile(path, dtype=np.uint8)
    xy, off, _bbox, _rings = nfx.geojson_pack(text, threads)
    return xy, off


def load_input_image(path):
    """load_input_image (src/main.rs:20-35): png/jpg/jpeg -> [H,W,3] u8."""
    ext = os.path.splitext(path)[1].lstrip(".")
    if ext in ("svs", "tif", "tiff"):     # InputImage::Slide (main.rs:21-24): the file bytes; level 0 is decoded on the GPU
        return np.fromfile(path, dtype=np.uint8)
    if ext not in ("png", "jpg", "jpeg"):
        die("Unsupported input format. Please use one of the following : svs, png, jpg, jpeg")
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8))


def extract_multi_gpu(image, xy, off, mask, gpus, patch_size, batch_size):
    """par_chunks over the GPUs: contiguous index ranges aligned to batch_size, one context per GPU
    (one host thread each, like the reference's rayon workers), merged in input order."""
    import nfx
    n = len(off) - 1
    F = len(nfx.feature_names(mask))
    cents = np.zeros((n, 2), np.float32)
    feats = np.zeros((n, F), np.float32)
    bounds = nfx.partition(n, batch_size, len(gpus))
    errors = []

    def work(k, gpu):
        lo, hi = bounds[k], bounds[k + 1]
        if hi <= lo:
            return
        try:
            with nfx.Extractor(gpu, patch_size, batch_size) as ex:
                if image.ndim == 1:
                    ex.load_tiff(image)
                else:
                    ex.upload_tile(image)
                ex.upload_polygons(xy[off[lo]:off[hi]], off[lo:hi + 1] - off[lo])
                ex.compute(mask)
                ex.download(cents[lo:hi], feats[lo:hi])
            log.info("Extracted features for %d/%d patches", hi, n)                 # main.rs:152-157
        except Exception as e:      # noqa: BLE001 -- reported below like the reference's panic
            errors.append(e)

    ts = [threading.Thread(target=work, args=(k, g)) for k, g in enumerate(gpus)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    return cents, feats


def write_csv_device(path, cents, feats, mask, device, rows_per_call=65536):
    """polars' CsvWriter output (oracle/SPEC.md B12) with the cells formatted on the GPU (nfx_csv_format)."""
    import nfx
    with open(path, "wb") as f, nfx.Extractor(device) as ex:
        f.write(nfx.csv_header(mask))
        for lo in range(0, len(cents), rows_per_call):
            f.write(ex.csv_format(cents[lo:lo + rows_per_call], feats[lo:lo + rows_per_call]))


def write_output(path, ext, keys, feats, names):
    """The writers of src/main.rs:160-189: first column `centroid` (Utf8), then one f32 column per feature."""
    import pyarrow as pa
    cols = {"centroid": pa.array(keys, type=pa.string())}
    for j, nm in enumerate(names):
        cols[nm] = pa.array(feats[:, j], type=pa.float32())
    table = pa.table(cols)
    if ext == "csv":
        import pyarrow.csv as pacsv
        pacsv.write_csv(table, path)
    elif ext in ("parquet", "pqt"):
        import pyarrow.parquet as pq
        pq.write_table(table, path)
    elif ext == "json":                       # polars JsonWriter default: one JSON object per line
        with open(path, "w") as f:
            for i in range(len(keys)):
                row = {"centroid": keys[i]}
                row.update({nm: (None if np.isnan(v) else float(v)) for nm, v in zip(names, feats[i])})
                f.write(json.dumps(row) + "\n")
    else:                                     # ipc | feather
        import pyarrow.feather as feather
        feather.write_feather(table, path, compression="uncompressed")


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if a.verbose else logging.INFO, format="%(levelname)s %(message)s")
    if a.verbose:
        print("Called Args :")
        print(a)
    ext = validate_paths(a)
    import nfx
    try:
        mask = nfx.parse_feature_sets(a.feature_sets)
    except nfx.NfxError as e:
        die(str(e))
    gpus = a.gpus if a.gpus else [0]
    log.info("Loading the geojson")
    xy, off = load_geometry(a.geometry)
    image = load_input_image(a.slide)
    log.info("Extracting features")
    try:
        cents, feats = extract_multi_gpu(image, xy, off, mask, gpus, a.patch_size, a.batch_size)
    except nfx.NfxError as e:
        die(str(e))
    keys = [nfx.centroid_key(c[0], c[1]) for c in cents]
    if ext == "csv":
        write_csv_device(a.output, cents, feats, mask, gpus[0])
    else:
        write_output(a.output, ext, keys, feats, nfx.feature_names(mask))
    return 0


if __name__ == "__main__":
    sys.exit(main())
