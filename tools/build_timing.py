"""Scene build times: python tools/build_timing.py [nx nz]   (default 64 32 = the C5 scene, 10,174,465 leaves).
Builds the bunny field with the device build and with the host build (RTP_DEVICE_BUILD=0), prints the phase times the library
reports (RTP_BUILD_TIMING) and the wall time of rtp_scene_create, and (below 3 M leaves) checks that both builds have equal digests."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTP_BUILD_TIMING"] = "1"

from rtp_b200 import api, scenes


def main():
    nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 32)
    api.init(0)
    t0 = time.perf_counter()
    sc = scenes.bunny_field(nx, nz)
    print(f"scene description ({len(sc.hittables)} hittables) made in {time.perf_counter() - t0:.2f} s", flush=True)
    digests = {}
    for mode in ("1", "0", "1"):
        os.environ["RTP_DEVICE_BUILD"] = mode
        t0 = time.perf_counter()
        g = api.Scene(sc)
        dt = time.perf_counter() - t0
        info = g.info()
        print(f"RTP_DEVICE_BUILD={mode}: rtp_scene_create {dt:.2f} s; {info.n_leaves} leaves, culling depth {info.culling_depth}, {info.device_bytes / 2**30:.2f} GiB on the device", flush=True)
        if len(sc.hittables) <= 3_000_000:
            digests[mode] = g.digest()
        g.close()
    if digests:
        print("digests equal:", digests["1"] == digests["0"], digests)


if __name__ == "__main__":
    main()
