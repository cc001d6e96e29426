"""Developer aid: end-to-end rate of msdr_chain_update (pinned host buffers) at C3 against the host chunk shape."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
m = importlib.import_module("minimal-sdr_b200")
import torch

K = m.load_ref_constants()
w = m.workloads.get("c3", K)
C, nb = 4096, 3446
hin, hout = m.capi.PinnedBuffer((C, nb * 128)), m.capi.PinnedBuffer((C, nb * 128))
hin.array[:] = np.random.default_rng(1).integers(-20000, 20000, hin.array.shape, dtype=np.int16)
for cc, nbk in ((0, 0), (4096, 16), (4096, 32), (4096, 128), (2048, 64), (2048, 128), (1024, 256), (4096, 24)):
    g = m.ReceiveChain(C)
    w.configure(g)
    g.set_option("host_chunk_channels", cc); g.set_option("host_chunk_blocks", nbk)
    g.update(hin.array, out=hout.array)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2):
        g.update(hin.array, out=hout.array)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"chunk channels {cc or 'auto':>5} blocks {nbk or 'auto':>5}: {2 * C * nb * 128 / dt / 1e9:6.2f} Gsamples/s = {2 * C * nb * 128 * 2 / dt / 1e9:5.1f} GB/s each way", flush=True)
    g.close()
