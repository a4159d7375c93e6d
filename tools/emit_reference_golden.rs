//! emit_reference_golden.rs -- dumps the REFERENCE's own results (oxabz/nuclei-feature-extraction with its pinned
//! tch-utils@d1c10c0 / geometric-features@163ae81 / tch 0.11) for the inputs of tests/golden/make_golden.py, so that the
//! rules this repo could only restate from their call sites (oracle/SPEC.md section B: "parity unpinned") can be pinned by
//! anyone who has cargo + libtorch. NOT BUILT in this repo's image (no rustc); it only uses calls that appear verbatim in
//! the reference (cited below), so it should compile as is next to them.
//!
//! How to run (in a checkout of the reference):
//!   1. python tests/golden/export_reference_inputs.py            (this repo) -> tests/golden/reference/case.{geojson,png}
//!   2. cp tools/emit_reference_golden.rs <reference>/src/bin/emit_reference_golden.rs
//!      and add to <reference>/Cargo.toml:
//!          [[bin]]
//!          name = "emit-reference-golden"
//!          path = "src/bin/emit_reference_golden.rs"
//!   3. cargo run --release --bin emit-reference-golden -- case.geojson case.png OUTDIR 64 20
//!   4. copy OUTDIR/reference_*.npy and reference_names.txt into this repo's tests/golden/reference/ and run
//!      python -m pytest tests/test_reference_golden.py      (skipped while the files are absent)
//!
//! Every array is written as a little-endian .npy (version 1.0), C order.
#![allow(dead_code)]

#[path = "../geojson.rs"]
mod geojson; // src/geojson.rs:8-24
#[path = "../utils.rs"]
mod utils; // src/utils.rs (preprocess_polygon, load_image_dataset, key strings)
#[path = "../features/mod.rs"]
mod features; // src/features/{mod,shape,color,texture}.rs

use std::io::Write;
use std::sync::{Arc, Mutex};

use polars::prelude::*;
use tch::{index::*, Device, Kind, Tensor};
use tch_utils::{
    glcm::glcm,
    glrlm::glrlm,
};

use features::FeatureSet;
use utils::PointsExt;

fn write_npy(path: &std::path::Path, descr: &str, shape: &[i64], bytes: &[u8]) {
    let dims = shape.iter().map(|d| d.to_string()).collect::<Vec<_>>().join(", ");
    let tuple = if shape.len() == 1 { format!("({},)", dims) } else { format!("({})", dims) };
    let mut header = format!("{{'descr': '{}', 'fortran_order': False, 'shape': {}, }}", descr, tuple);
    while (10 + header.len() + 1) % 64 != 0 {
        header.push(' ');
    }
    header.push('\n');
    let mut f = std::fs::File::create(path).expect("create npy");
    f.write_all(b"\x93NUMPY\x01\x00").unwrap();
    f.write_all(&(header.len() as u16).to_le_bytes()).unwrap();
    f.write_all(header.as_bytes()).unwrap();
    f.write_all(bytes).unwrap();
}

fn dump_f32(dir: &std::path::Path, name: &str, t: &Tensor) {
    let shape = t.size();
    let v = Vec::<f32>::from(t.to_kind(Kind::Float).to_device(Device::Cpu).contiguous().view([-1]));
    let bytes: Vec<u8> = v.iter().flat_map(|x| x.to_le_bytes()).collect();
    write_npy(&dir.join(format!("reference_{}.npy", name)), "<f4", &shape, &bytes);
}

fn dump_f64s(dir: &std::path::Path, name: &str, rows: usize, cols: usize, v: &[f64]) {
    let bytes: Vec<u8> = v.iter().flat_map(|x| x.to_le_bytes()).collect();
    write_npy(&dir.join(format!("reference_{}.npy", name)), "<f8", &[rows as i64, cols as i64], &bytes);
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    assert!(a.len() >= 6, "usage: emit-reference-golden <geojson> <png> <outdir> <patch_size> <batch_size>");
    let out = std::path::PathBuf::from(&a[3]);
    std::fs::create_dir_all(&out).unwrap();
    let patch_size: usize = a[4].parse().unwrap();
    let batch_size: usize = a[5].parse().unwrap();
    let _ = tch::no_grad_guard();

    // src/main.rs:37-42 (load_geometry) and src/main.rs:20-35 (load_input_image, image branch)
    let file = std::fs::File::open(&a[1]).unwrap();
    let geometry: geojson::FeatureCollection = serde_json::from_reader(std::io::BufReader::new(file)).unwrap();
    let image = Arc::new(Mutex::new(tch::vision::image::load(&a[2]).unwrap()));
    let n = geometry.features.len();

    let sets: Vec<(&str, Box<dyn FeatureSet>)> = vec![
        ("geometry", Box::new(features::ShapeFeatureSet)),
        ("color", Box::new(features::ColorFeatureSet)),
        ("glcm", Box::new(features::GlcmFeatureSet)),
        ("glrlm", Box::new(features::GLRLMFeatureSet)),
        ("gabor", Box::new(features::GaborFilterFeatureSet)),
    ];
    let mut names: Vec<String> = Vec::new();
    let mut columns: Vec<Vec<f64>> = Vec::new(); // [column][row]
    let mut keys: Vec<String> = Vec::new();
    let mut all_masks: Vec<Tensor> = Vec::new();
    let mut all_patches: Vec<Tensor> = Vec::new();
    let mut poly_geom: Vec<f64> = Vec::new();

    for (ci, chunk) in geometry.features.chunks(batch_size).enumerate() {
        // src/main.rs:149 -> src/input.rs:13-30 -> src/utils.rs:141-206
        let (centroids, polygons, patches, masks) = utils::load_image_dataset(chunk, image.clone(), patch_size);
        keys.extend(utils::centroids_to_key_strings(&centroids));
        let mut col = 0usize;
        for (_set, fs) in &sets {
            // src/main.rs:47-91: every set on the same batch
            let df = fs.compute_features_batched(&centroids, &polygons, &patches, &masks);
            for s in df.get_columns() {
                if s.name() == "centroid" {
                    continue;
                }
                let v: Vec<f64> = s.cast(&DataType::Float64).unwrap().f64().unwrap().into_iter().map(|x| x.unwrap_or(f64::NAN)).collect();
                if ci == 0 {
                    names.push(s.name().to_string());
                    columns.push(Vec::new());
                }
                columns[col].extend(v);
                col += 1;
            }
        }
        // un-vendored rules, one tap each (call sites cited)
        for poly in &polygons {
            let p = poly.to_tchutils_points(); // src/features/shape.rs:67-68
            let hull = geometric_features::convex_hull::convex_hull_features(&p); // shape.rs:93-97
            poly_geom.extend([
                geometric_features::area(&p) as f64,                 // shape.rs:89
                geometric_features::perimeter(&p) as f64,            // shape.rs:90
                geometric_features::equivalent_perimeter(&p) as f64, // shape.rs:91
                geometric_features::compacity(&p) as f64,            // shape.rs:92
                hull.area as f64,
                hull.perimeter as f64,
                hull.deviation as f64,
            ]);
        }
        if ci == 0 {
            let take = 4.min(patches.size()[0]);
            let p4 = patches.i(..take);
            let m4 = masks.i(..take);
            dump_f32(&out, "hsv", &tch_utils::color::hsv_from_rgb(&p4)); // color.rs:45
            dump_f32(&out, "hed", &tch_utils::color::hed_from_rgb(&p4)); // color.rs:46
            let gs = p4.mean_dim(Some(&([-3][..])), true, Kind::Float); // texture.rs:36
            for (lv, off, tag) in [(32u8, (0i64, 1i64), "32_0_1"), (64, (1, 1), "64_1_1"), (128, (1, 0), "128_1_0"), (254, (1, -1), "254_1_-1"), (254, (0, 1), "254_0_1")] {
                dump_f32(&out, &format!("glcm_{}", tag), &glcm(&gs, off, lv, Some(&m4), true)); // texture.rs:40-46
            }
            for (d, tag) in [((1i64, 0i64), "1_0"), ((1, 1), "1_1"), ((0, 1), "0_1"), ((-1, 1), "-1_1")] {
                dump_f32(&out, &format!("glrlm_{}", tag), &glrlm(&gs, 24, 16, d, Some(&m4)).to_kind(Kind::Float)); // texture.rs:193-194
            }
            let g2 = gs.i(..2.min(take));
            dump_f32(&out, "gabor", &tch_utils::gabor::apply_gabor_filter(&g2, 8, 30, &[0.5, 1.0, 2.0, 4.0, 6.0, 8.0], 0.45)); // texture.rs:333-334
            // the ellipse rule on fixed parameters (shape.rs:80-87)
            let e = tch_utils::shapes::ellipse(patch_size, patch_size, (1.25, -2.5), (17.0, 9.5), 0.6, (Kind::Float, Device::Cpu));
            dump_f32(&out, "ellipse_fixed", &e);
        }
        all_masks.push(masks);
        all_patches.push(patches);
    }
    dump_f32(&out, "masks", &Tensor::cat(&all_masks, 0));     // tch_utils::shapes::polygon, utils.rs:152-157
    dump_f32(&out, "patches", &Tensor::cat(&all_patches, 0)); // utils.rs:159-192
    let f = names.len();
    let mut flat = vec![0f64; n * f];
    for (c, colv) in columns.iter().enumerate() {
        for (r, v) in colv.iter().enumerate() {
            flat[r * f + c] = *v;
        }
    }
    dump_f64s(&out, "features", n, f, &flat);
    dump_f64s(&out, "polygon_geometry", n, 7, &poly_geom);
    std::fs::write(out.join("reference_names.txt"), names.join("\n")).unwrap();
    std::fs::write(out.join("reference_keys.txt"), keys.join("\n")).unwrap();
    println!("wrote {} nuclei x {} columns to {}", n, f, out.display());
}
