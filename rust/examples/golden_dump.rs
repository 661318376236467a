//! Reference-side pin for the CPU oracle of the B200 path. Uses only the crate's public items, unmodified:
//! `obj::load` (mesh.rs:145), `Bvh::new` (bvh.rs:70), `Hittable::hit` (hittable.rs:18), `Camera::shoot` (render.rs:32),
//! `Transformation::lookat` (utility.rs:172). Run from the crate root (it reads assets/bunny.obj):
//!
//!     cargo run --release --example golden_dump
//!
//! Output `c2_hits_every97.ref.bin`, little endian, one 48-byte record per ray k = 0, 97, 194, ... of the 1920x1080 pixel-centre
//! batch (i fastest, row j = 0 at the bottom, u = (i + 0.5) / W, v = (j + 0.5) / H):
//!     u64 t_bits (f64::to_bits of Hit::t; +inf on a miss)   u32 material (0xFFFFFFFF on a miss)   u32 n_candidates
//!     u32 candidates[8]  (indices into the Vec<Hittable> handed to Bvh::new whose own `hit` returns exactly this t; unused = 0xFFFFFFFF)
use raytracing2::bvh::Bvh;
use raytracing2::hittable::Hittable;
use raytracing2::material::{Absorb, Emit, Material, MaterialId, Scatter};
use raytracing2::mesh::{obj, MeshId};
use raytracing2::randomness::Randomizer;
use raytracing2::render::{Camera, SceneData};
use raytracing2::utility::*;
use nalgebra::vector;
use rand::SeedableRng;
use std::f64::consts::FRAC_PI_4;
use std::io::Write;

fn main() {
    let (w, h) = (1920u32, 1080u32);
    // example_scenes.rs:309-350 `bunny()`; materials and the sky do not influence closest hits
    let bunny = obj::load("assets/bunny.obj").unwrap();
    let mut list: Vec<Hittable> = bunny.iter_triangles().map(|tid| Hittable::Triangle { triangle: tid, mesh: MeshId(0) }).collect();
    list.push(Hittable::Sphere { center: vector![0.0, -1000.0, -1.0], radius: 1000.0, material: MaterialId(1) });
    let scene_data = SceneData {
        material_table: vec![
            Material::new(Scatter::Lambert, Absorb::Albedo(rgb(0.8, 0.8, 0.8)), Emit::None),
            Material::new(Scatter::Metal { fuzziness: 0.05 }, Absorb::Albedo(rgb(0.8, 0.8, 0.8)), Emit::None),
        ],
        mesh_table: vec![bunny],
        texture_table: vec![],
    };
    let root = Hittable::Bvh(Bvh::new(list.clone(), &scene_data));
    let camera = Camera {
        aspect_ratio: w as Real / h as Real, // main.rs:22
        fov: FRAC_PI_4,
        focal_dist: 1.0,
        lens_radius: 0.0,
        transformation: Transformation::lookat(&vector![-1.5, 1.5, 2.5], &vector![0.0, 0.5, 0.0], &vector![0.0, 1.0, 0.0]),
    };
    let mut rng = Randomizer::seed_from_u64(0); // lens_radius = 0: the lens draw (render.rs:36) does not reach the ray
    let mut out = std::fs::File::create("c2_hits_every97.ref.bin").unwrap();
    let mut n = 0usize;
    let mut k = 0u64;
    while k < (w as u64) * (h as u64) {
        let (i, j) = ((k % w as u64) as Real, (k / w as u64) as Real);
        let uv = vector![(i + 0.5) / w as Real, (j + 0.5) / h as Real];
        let ray = camera.shoot(uv, &mut rng);
        let (t_bits, material, cands) = match root.hit(&ray, &scene_data) {
            None => (f64::INFINITY.to_bits(), u32::MAX, Vec::new()),
            Some((hit, material)) => {
                let mut c = Vec::new();
                for (id, item) in list.iter().enumerate() {
                    if let Some((hh, _)) = item.hit(&ray, &scene_data) {
                        if hh.t.to_bits() == hit.t.to_bits() { c.push(id as u32); }
                    }
                }
                (hit.t.to_bits(), material.0, c)
            }
        };
        out.write_all(&t_bits.to_le_bytes()).unwrap();
        out.write_all(&material.to_le_bytes()).unwrap();
        out.write_all(&(cands.len() as u32).to_le_bytes()).unwrap();
        for s in 0..8 {
            let v = if s < cands.len() { cands[s] } else { u32::MAX };
            out.write_all(&v.to_le_bytes()).unwrap();
        }
        n += 1;
        k += 97;
    }
    println!("wrote c2_hits_every97.ref.bin: {} records", n);
}
