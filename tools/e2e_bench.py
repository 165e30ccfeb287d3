"""Host-buffer path (FootsiesEnv.step_host -> fg_step_host_compact) at the bench workload for several slice sizes, next
to the device-layout call fg_step_host.  usage: python tools/e2e_bench.py [num_envs] [steps]"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n, steps, layout):
    import torch
    from footsies_gym_b200 import FootsiesEnv, _capi
    env = FootsiesEnv(num_envs=n, seed=0)
    env.reset()
    g = torch.Generator().manual_seed(1)
    tapes = [torch.randint(0, 8, (n,), generator=g, dtype=torch.uint8).pin_memory() for _ in range(4)]
    for i in range(100):                                   # desynchronise the episodes a little
        env.step(tapes[i % 4].cuda())
    if layout == "f32":
        pin = dict(pin_memory=True)
        bufs = [torch.zeros((n, 8), **pin), torch.zeros(n, **pin), torch.zeros(n, dtype=torch.uint8, **pin),
                torch.zeros(n, dtype=torch.int32, **pin), torch.zeros((n, 4), dtype=torch.uint8, **pin)]

        def one(i):
            _capi.check(env._lib.fg_step_host(env._handle, C.c_void_p(tapes[i % 4].data_ptr()), None,
                                              *[C.c_void_p(b.data_ptr()) for b in bufs], None))
    elif layout == "packed":
        def one(i):
            env.step_host_packed(tapes[i % 4])
    else:
        def one(i):
            env.step_host(tapes[i % 4])
    for i in range(3):
        one(i)
    f0 = env.episode_stats()["env_frames"]
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    frames = env.episode_stats()["env_frames"] - f0
    h2d, d2h = env.host_io_bytes_per_step()
    if layout == "f32":
        d2h = n * 45
    if layout == "packed":
        h2d, d2h = env.host_io_bytes_per_step(packed=True)
    print(json.dumps({"layout": layout, "chunk_envs": os.environ.get("FOOTSIES_B200_HOST_CHUNK_ENVS", "default"),
                      "lead_div": os.environ.get("FOOTSIES_B200_HOST_LEAD_DIV", "default"),
                      "env_frames_per_sec": frames / dt, "ms_per_step": dt / steps * 1e3,
                      "d2h_GBps": d2h * steps / dt / 1e9}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--lead":
        # packed layout: first slice = chunk / div (FOOTSIES_B200_HOST_LEAD_DIV), interleaved repetitions
        for rep in range(3):
            for div in (1, 2, 4):
                env = dict(os.environ, FOOTSIES_B200_HOST_LEAD_DIV=str(div))
                subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(4 * 1024 * 1024), "60", "packed"], env=env, check=False)
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    for layout, chunk in [("f32", 1 << 30), ("f32", 512 * 1024), ("compact", 1 << 30), ("compact", 2 * 1024 * 1024),
                          ("compact", 1024 * 1024), ("compact", 512 * 1024), ("compact", 256 * 1024), ("compact", 128 * 1024)]:
        env = dict(os.environ, FOOTSIES_B200_HOST_CHUNK_ENVS=str(chunk))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n), str(steps), layout], env=env, check=False)
