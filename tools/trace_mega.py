"""Per-phase timeline of the persistent decode kernel (LIVECC_B200_MEGA_TRACE=1): consumer thread 0 of every CTA stamps
%globaltimer after RMSNorm staging (A), after its last tile + epilogue (B) and after the grid barrier (C) of every phase of
the first layers. Prints, per phase, the median / max over CTAs of stage, work, barrier-wait and the phase's wall time."""
import os
import sys

os.environ["LIVECC_B200_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from livecc_b200 import _cabi
from livecc_b200.config import LiveCCConfig
from livecc_b200.engine import LiveCCB200ForConditionalGeneration

cfg = LiveCCConfig.livecc_7b()
eng = LiveCCB200ForConditionalGeneration.from_synthetic(cfg, device="cuda")
L = cfg.text_config.num_hidden_layers
kv = int(os.environ.get("KV", "8000"))
g = torch.Generator().manual_seed(0)
out = eng.generate_batch([dict(input_ids=torch.randint(1000, 9000, (1, kv), generator=g).cuda())], max_new_tokens=1)
cache = out[0].past_key_values
with torch.inference_mode():
    cache.scalars[_cabi.SC_FINISHED] = 0
st = [cache.stream_state()]
for _ in range(3):
    eng._native.decode_mega_debug(st, 0, L, 31, 1)
torch.cuda.synchronize()
off = eng._native.mega_trace_offset
G = eng.ctx.num_sms
tr = eng._native.workspace[off:off + 256 * 64 * 8].view(torch.int64).view(256, 64)[:G].cpu().double() / 1e3   # us
t0 = tr[:, 0].min()
tr = tr - t0
names = [("qkv", 3), ("attn", 2), ("o_proj", 3), ("gate_up", 3), ("down", 2)]
idx = 1
prev_end = tr[:, 0]
for layer in range(4):
    for name, n in names:
        if idx + n > 64:
            break
        stamps = tr[:, idx:idx + n]
        idx += n
        if n == 3:
            stage = stamps[:, 0] - prev_end
            work = stamps[:, 1] - stamps[:, 0]
        else:
            stage = torch.zeros(G, dtype=torch.float64)
            work = stamps[:, 0] - prev_end
        wait = stamps[:, -1] - stamps[:, -2]
        wall = stamps[:, -1].max() - prev_end.min()
        print(f"layer {layer} {name:8s} wall {wall:6.2f} us | stage med {stage.median():5.2f} max {stage.max():5.2f} | "
              f"work med {work.median():6.2f} max {work.max():6.2f} min {work.min():6.2f} | barrier wait med {wait.median():5.2f} "
              f"min {wait.min():5.2f} | exit spread {stamps[:, -1].max() - stamps[:, -1].min():4.2f}")
        prev_end = stamps[:, -1]
print("kv_len", kv, "err", eng._native.mega_error())
