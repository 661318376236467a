"""The reference's `main.rs` on the B200 core, step for step (main.rs:12-131): pick a scene, 800x600, 4 spp, depth 8,
32x32 tiles, render, merge, to_srgb_u8, write output.tga.

    python examples/main.py [scene] [--tiles] [--gpus N] [--out output.tga]

Default: one rtp_render_srgb8 call for the whole frame (output stage on the device). `--tiles` keeps the reference's work
partition literally — one rtp_render call per 32x32 tile in the job queue's order (LIFO, main.rs:58) and the merge +
to_srgb_u8 on the host (main.rs:110-122) — and produces the same bytes, because a pixel's samples are keyed by (pixel,
sample) and not by the tile or the thread that traced them. `--gpus N` is the reference's `num_workers` (main.rs:27) on this
core: the scene is replicated on N GPUs (rtp_scene_create_multi) and the SAME single call renders on all of them — again the same bytes."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from rtp_b200 import api, scenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", nargs="?", default="bunny", choices=["three_balls", "more_balls", "more_balls_optimized", "two_balls", "earth", "one_triangle", "glass_bunny", "bunny", "bunny_lambert", "demo"])
    ap.add_argument("--tiles", action="store_true")
    ap.add_argument("--out", default="output.tga")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()

    output_width, output_height = 800, 600          # main.rs:13
    example_scene = getattr(scenes, args.scene)()    # main.rs:16-21
    max_bounce, num_samples = 8, 4                   # main.rs:25, 32
    tile_w = tile_h = 32                             # main.rs:36
    api.init(0)
    scene = api.Scene(example_scene, device_mask=(1 << args.gpus) - 1 if args.gpus > 1 else 0)  # Bvh::new happens in here

    t0 = time.perf_counter()
    if args.tiles:
        jobs = [tuple(int(x) for x in t) for t in api.split_in_tiles(output_width, output_height, tile_w, tile_h)]
        frame = np.zeros((output_height, output_width, 3))
        while jobs:
            tile = jobs.pop()                        # main.rs:58: the queue is popped from the end
            scene.render(output_width, output_height, num_samples, max_bounce, seed=args.seed, tile=tile, out=frame, foreground=False)
        rgba = api.to_srgb_u8(frame)                 # main.rs:110-122
    else:
        rgba, _ = scene.render_srgb8(output_width, output_height, num_samples, max_bounce, seed=args.seed)
    dt = time.perf_counter() - t0
    print(f"Rendering done in {dt:.4f} seconds")     # main.rs:106
    api.tga.save(rgba, args.out)                     # main.rs:125-126
    return rgba


if __name__ == "__main__":
    main()
