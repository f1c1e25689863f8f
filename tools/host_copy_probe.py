"""Where does the host side of the end-to-end path spend its time?  Copy rates page cache (mmap) -> pinned staging
with 1..N threads, cudaHostRegister cost per GB, H2D rate from pinned / registered memory."""
import ctypes, mmap, os, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "statdyn-analysis_b200"))
import numpy as np
import torch
from sdanalysis_b200 import _lib
lib = _lib.load()
size = 28_000_000
path = "/dev/shm/sdb_probe.bin"
n_frames = 32
with open(path, "wb") as f:
    f.write(np.random.default_rng(0).integers(0, 255, size * n_frames, dtype=np.uint8).tobytes())
fh = open(path, "r+b"); mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_WRITE)
src = np.frombuffer(mm, dtype=np.uint8)
pinned = torch.empty(size, dtype=torch.uint8).pin_memory()
dst_ptr = pinned.data_ptr()
plain = np.empty(size, dtype=np.uint8)
dev = torch.empty(size, dtype=torch.uint8, device="cuda")
def rate(fn, reps=n_frames):
    t = time.perf_counter()
    for i in range(reps): fn(i)
    return size * reps / (time.perf_counter() - t) / 1e9
print("first touch (page faults) numpy copy mmap->plain GB/s: %.1f" % rate(lambda i: np.copyto(plain, src[i*size:(i+1)*size])))
print("numpy copy mmap->plain GB/s: %.1f" % rate(lambda i: np.copyto(plain, src[i*size:(i+1)*size])))
print("numpy copy mmap->pinned GB/s: %.1f" % rate(lambda i: np.copyto(pinned.numpy(), src[i*size:(i+1)*size])))
print("numpy copy plain->plain GB/s: %.1f" % rate(lambda i: np.copyto(plain, plain[::1] if False else pinned.numpy())))
base = src.ctypes.data
print("engine copy pool (SDB_COPY_THREADS=%s) mmap->pinned GB/s: %.1f" % (os.environ.get("SDB_COPY_THREADS", "default"), rate(lambda i: lib.sdb_test_host_copy(dst_ptr, base + i * size, size))))
torch.cuda.synchronize()
def h2d(i):
    dev.copy_(pinned, non_blocking=True)
t = time.perf_counter(); [h2d(i) for i in range(n_frames)]; torch.cuda.synchronize()
print("H2D pinned GB/s: %.1f" % (size * n_frames / (time.perf_counter() - t) / 1e9))
t = time.perf_counter(); rc = lib.sdb_host_register(base, size * n_frames); dt = time.perf_counter() - t
print("cudaHostRegister rc %d: %.3f s for %.2f GB = %.2f GB/s" % (rc, dt, size * n_frames / 1e9, size * n_frames / dt / 1e9))
if rc == 0:
    cudart = torch.cuda.cudart()
    reg = torch.from_numpy(src)
    t = time.perf_counter()
    for i in range(n_frames):
        dev.copy_(reg[i*size:(i+1)*size], non_blocking=True)
    torch.cuda.synchronize()
    print("H2D registered mmap GB/s: %.1f" % (size * n_frames / (time.perf_counter() - t) / 1e9))
    t = time.perf_counter(); lib.sdb_host_unregister(base); print("unregister %.3f s" % (time.perf_counter() - t))
os.unlink(path)
