cd $GRAFT_REPO_ROOT
for n in 512 1024; do GTOK_EXP_WALKL_BIG=1024 timeout 300 python - <<PY 2>&1 | tail -2
import sys, os, ctypes as C
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "glearning-benchmark_b200")]
import torch
from gtok import engine, synth
from gtok.engine import lib, _ptr, _stream, _csr_c, _sent_c, check
dev = torch.device("cuda:0")
N = $n
G = 32_000_000 // N
gb = synth.er_sparse(G, N, mean_degree=8.0, seed=N, device=dev)
csr = engine.build_csr(gb)
cfg = engine.SentConfig(idx_offset=6, max_length=60000)
sc, ch = _sent_c(cfg, 0, 0, None, None)
cc = _csr_c(csr)
n_total = int(csr.rowptr.numel()) - 1
nbytes = n_total * 12 + 16
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
ln = torch.empty(G, dtype=torch.int32, device=dev)
print("log bytes api", lib.gtok_sent_log_bytes(C.byref(cc), C.byref(sc), n_total))
def walk():
    check(lib.gtok_sent_walk(C.byref(cc), C.byref(sc), C.byref(ch), n_total, _ptr(ws), nbytes, _ptr(ln), _stream()))
for _ in range(2): walk()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): walk()
b.record(); torch.cuda.synchronize()
print(f"N={N} G={G} list walk (thread per graph): {a.elapsed_time(b)/3:.3f} ms  mean_len {float(ln.float().mean()):.0f}")
PY
done
