"""One invocation of every kernel class on its benchmark shape, for `ncu --set full -k regex:...` captures (profiles/).
Each section launches its kernel(s) twice (first = warm-up; capture the second with --launch-skip / -c as needed).
Shapes: the 4B encoder layer at M = 64 x 197 = 12608 rows (BASELINE configs[1]); InfoNCE head 1024 x 8192 x 1536 (configs[3]);
fused Adam on 200 M bf16 parameters with fp32 master (28 B / parameter)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from one_peace_b200 import kernels as K, relpos
from one_peace_b200.one_peace import OnePeaceRetrievalConfig, OnePeaceRetrievalModel
from one_peace_b200.unify_model_config import one_peace_4b_encoder_config
from one_peace_b200.optim.adam import Adam
from one_peace_b200.criterions.image_text_retrieval_loss import itc_loss
import synth

dev = torch.device("cuda")
which = set(sys.argv[1:]) or {"layer", "adam", "ln", "embed", "infonce", "audio"}
bf = torch.bfloat16
if "layer" in which:
    cfg = OnePeaceRetrievalConfig()
    cfg.encoder = one_peace_4b_encoder_config(layers=2, embed_dim=1536, ffn_embed_dim=6144, attention_heads=24, patch_image_size=224)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OnePeaceRetrievalModel(cfg, None, "image")
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "gamma_" in n: p.fill_(0.1)
                elif "rel_pos_table" in n: p.normal_(0, 0.1)
    model.eval()
    img = torch.randn(64, 3, 224, 224, device=dev)
    with torch.no_grad():
        for _ in range(2):
            model(src_images=img, encoder_type="image")
    torch.cuda.synchronize()
if "adam" in which:
    n = 200_000_000
    p = torch.nn.Parameter(torch.zeros(n, dtype=bf, device=dev)); p.grad = torch.full((n,), 1e-3, dtype=bf, device=dev)
    opt = Adam([p], lr=1e-4, betas=(0.9, 0.98), weight_decay=0.05, master_weights=True)
    for _ in range(2):
        opt.grad_norm_and_scale(1.0, 3.0); opt.step()
    torch.cuda.synchronize(); del opt, p
if "ln" in which:
    rows, d = 4 * 12608, 1536
    x = torch.randn(rows, d, device=dev); w, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    y = torch.empty(rows, d, dtype=bf, device=dev); dy = torch.randn(rows, d, device=dev).to(bf); dx = torch.zeros(rows, d, device=dev)
    dg, db = torch.empty(d, device=dev), torch.empty(d, device=dev)
    for _ in range(2):
        K.layernorm(x, w, b, y); K.layernorm_bwd(x, dy, w, b, dx, accumulate=True, dgamma=dg, dbeta=db)
    u = torch.randn(rows // 2, 2 * 6144, device=dev).to(bf); uo = torch.empty(rows // 2, 6144, dtype=bf, device=dev)
    for _ in range(2):
        K.geglu_fwd(u, uo)
    torch.cuda.synchronize()
if "embed" in which:
    tok = torch.randint(4, 50264, (1024, 71), device=dev); table = torch.randn(50264, 1536, device=dev).to(bf)
    pos, cls = torch.randn(514, 1536, device=dev), torch.randn(1536, device=dev)
    for _ in range(2):
        K.text_embed(tok, table, pos, cls)
    torch.cuda.synchronize()
if "infonce" in which:
    a_all, t_all = synth.contrastive_pair(8192, 1536, seed=123)
    a = a_all[:1024].to(dev).requires_grad_(True); t = t_all[:1024].to(dev).requires_grad_(True)
    ls = torch.tensor(2.659, device=dev, requires_grad=True)
    ga, gt = a_all.to(dev), t_all.to(dev)
    for _ in range(2):
        loss, _, _ = itc_loss(a, t, ga, gt, ls.exp(), 0, 0.0); loss.backward()
    torch.cuda.synchronize()
if "audio" in which:
    Ba, N = 16, 240000
    wav = torch.randn(Ba, N, device=dev); frames = (N - 10) // 5 + 1
    a0 = torch.empty(Ba * frames, 16, dtype=bf, device=dev); w0 = torch.randn(512, 16, device=dev).to(bf)
    y0 = torch.empty(Ba * frames, 512, dtype=bf, device=dev)
    for _ in range(2):
        K.audio_frame10(wav, frames, a0); K.gemm(a0, w0, K.EPI_STORE_BF16, y0)
    torch.cuda.synchronize()
print("done", sorted(which))
