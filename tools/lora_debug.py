import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import cosine, load_golden
import test_lora_gpu as T
for name in ["clip_small", "siglip_small"]:
    fx = load_golden(f"tower_lora_{name}.pt")
    wrap = T._build(fx)
    x = T._norm_input(fx)
    cls, pc, pt5, out, loss = T._run(wrap, x)
    print(name, "loss", loss.item(), fx["loss"].item(), "lhs cos", cosine(out.last_hidden_state, fx["last_hidden_state"]))
    loss.backward()
    P = dict(wrap.model.named_parameters())
    for k, gref in fx["grads"].items():
        if k.startswith("project_"):
            seq, idx, leaf = k.split(".")
            g = getattr(getattr(wrap, seq)[int(idx)], leaf).grad
        elif k.endswith(("lora_A", "lora_B")):
            pair = wrap.model.lora[k.rsplit(".", 1)[0].replace(".", "/")]
            g = pair.A.grad if k.endswith("lora_A") else pair.B.grad
        else:
            g = P[k].grad
        print(f"  {k:70s} cos {cosine(g, gref):.4f} norm ratio {g.float().norm().item() / (gref.norm().item() + 1e-30):.4f}")
