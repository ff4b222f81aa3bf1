"""Launch-bound small-batch inference: 8 texts (16 tokens) through the 4B text branch (40 layers), eager launch sequence vs
one_peace_b200.graphs.GraphedForward replay (hub cuda_graph=True).  CUDA-event timed."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from one_peace_b200.one_peace import OnePeaceRetrievalConfig, OnePeaceRetrievalModel
from one_peace_b200.one_peace.hub_interface import OnePeaceHubInterface, _Dictionary
from one_peace_b200.unify_model_config import one_peace_4b_encoder_config
dev = torch.device("cuda")
cfg = OnePeaceRetrievalConfig()
cfg.encoder = one_peace_4b_encoder_config(layers=40, embed_dim=1536, ffn_embed_dim=6144, attention_heads=24, patch_image_size=224)
torch.manual_seed(0)
with torch.device(dev):
    model = OnePeaceRetrievalModel(cfg, _Dictionary(50264), "vl")
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma_" in n: p.fill_(0.1)
            elif "rel_pos_table" in n: p.normal_(0, 0.1)
model = model.to(torch.bfloat16).eval()
tok = torch.randint(4, 50264, (8, 16), device=dev)
tok[:, 0] = 0
res = {}
for name, graph in (("eager", False), ("cuda graph", True)):
    hub = OnePeaceHubInterface(model, device="cuda", cuda_graph=graph)
    for _ in range(3):
        out = hub.extract_text_features(tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = hub.extract_text_features(tok)
    e1.record(); torch.cuda.synchronize()
    res[name] = (e0.elapsed_time(e1) / 20, out.float().clone())
print(f"8 texts x 16 tokens, 4B text branch (40 layers): eager {res['eager'][0]:.3f} ms | cuda graph {res['cuda graph'][0]:.3f} ms | "
      f"identical outputs: {torch.equal(res['eager'][1], res['cuda graph'][1])}")
