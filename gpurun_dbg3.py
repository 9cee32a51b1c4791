import torch, sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import biomedkg_b200 as b
from biomedkg_b200.draws import ReplayDraws, set_draws
from conftest import rel_err
for name in ["grace_none","grace_mean2","grace_attention","dgi_none","ggd_none_a","ggd_none_b","ggd_redaf"]:
    fx=torch.load(f"tests/golden/{name}.pt",weights_only=False); cfg=fx["cfg"]
    mod=getattr(b,cfg["cls"])(in_dim=cfg["in_dim"],hidden_dim=cfg["hidden_dim"],out_dim=cfg["out_dim"],num_hidden_layers=cfg["num_hidden_layers"],fuse_method=cfg["fuse_method"])
    mod.load_state_dict({k:v.float() for k,v in fx["state_dict"].items()}); mod=mod.cuda(); mod.train(name!="ggd_redaf")
    set_draws(mod, ReplayDraws(fx["draws"],"cuda"))
    class B: x=fx["x"].float().cuda(); edge_index=fx["edge_index"].cuda()
    loss=mod.training_step(B); loss.backward()
    grads={k:p.grad for k,p in mod.named_parameters() if p.grad is not None}
    flat=torch.cat([grads[k].flatten().double().cpu() for k in sorted(grads)]); ref=torch.cat([fx["grads"][k].flatten().double() for k in sorted(grads)])
    cos=float((flat*ref).sum()/(flat.norm()*ref.norm()))
    print(name,"loss",float(loss),float(fx["loss"]),"flat rel",float((flat-ref).norm()/ref.norm()),"cos",cos)
    print("   ",{k.replace("model.encoder.graph_layers","L"):round(rel_err(g,fx["grads"][k]),4) for k,g in grads.items() if fx["grads"][k].norm()>1e-12})
    mod.eval()
    with torch.no_grad(): emb=mod(B.x,B.edge_index)
    print("    embed rel", rel_err(emb,fx["embed_eval"]))
