"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``local_3d_attention.py`` / ``vq.py`` / ``autoencoder.py`` by path from
``/root/reference/vq-video-diffusion`` (read-only), runs them on seeded CPU inputs in
fp32 and writes small ``.npz`` files.  Large inputs are not stored: they are
re-drawn from a seeded ``torch.Generator`` by :func:`seeded` and guarded by a
checksum stored in the fixture.  ``main.py`` / ``train_vqae.py`` cannot be imported
(matplotlib / MNIST download), so the denoiser wrapper is composed here from the
reference transformer plus a Linear head, which is all ``main.py:25-36`` does.
"""
import os
import sys

import numpy as np
import torch

REF = '/root/reference/vq-video-diffusion'
HERE = os.path.dirname(os.path.abspath(__file__))


def seeded(seed, *shape, kind='randn', hi=None):
    g = torch.Generator().manual_seed(seed)
    if kind == 'randn':
        return torch.randn(*shape, generator=g)
    if kind == 'rand':
        return torch.rand(*shape, generator=g)
    return torch.randint(0, hi, shape, generator=g)


def checksum(t):
    t = t.double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 977
    return np.array([t.sum().item(), (t * w).sum().item()])


def sd_np(module):
    return {'sd/' + k: v.detach().numpy() for k, v in module.state_dict().items()}


def main():
    sys.path.insert(0, REF)
    from local_3d_attention import Local3dAttention, Local3dAttentionTransformer  # noqa
    from vq import VectorQuantizerEMA  # noqa
    from autoencoder import SimpleResidualEncoder, SimpleResidualDecoder  # noqa
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- 1. small module: everything stored ---------------------------------------
    torch.manual_seed(1)
    ext = (1, 2, 2)
    m = Local3dAttention(ext, dim=32, heads=2, dim_head=16, use_checkpointing=False)
    x = torch.randn(1, 3, 6, 5, 32, requires_grad=True)
    q = torch.randn(1, 3, 6, 5, 32, requires_grad=True)
    dout = torch.randn(1, 3, 6, 5, 32)
    out = m(x, q)
    out.backward(dout)
    np.savez_compressed(os.path.join(HERE, 'attn_small.npz'), ext=np.array(ext), heads=2, dim_head=16,
                        x=x.detach().numpy(), q=q.detach().numpy(), dout=dout.numpy(), out=out.detach().numpy(),
                        dx=x.grad.numpy(), dq=q.grad.numpy(),
                        **sd_np(m), **{'grad/' + k: v.grad.numpy() for k, v in m.named_parameters()})

    # ---- 1b. heads==1 and dim_head==dim: to_out is Identity ------------------------
    torch.manual_seed(2)
    ext = (2, 1, 1)
    m = Local3dAttention(ext, dim=16, heads=1, dim_head=16, use_checkpointing=True)
    x = torch.randn(2, 5, 4, 4, 16, requires_grad=True)
    q = torch.randn(2, 5, 4, 4, 16, requires_grad=True)
    dout = torch.randn(2, 5, 4, 4, 16)
    out = m(x, q)
    out.backward(dout)
    np.savez_compressed(os.path.join(HERE, 'attn_noproj.npz'), ext=np.array(ext), heads=1, dim_head=16,
                        x=x.detach().numpy(), q=q.detach().numpy(), dout=dout.numpy(), out=out.detach().numpy(),
                        dx=x.grad.numpy(), dq=q.grad.numpy(), **sd_np(m))

    # ---- 2. config 1 (B2, 8x16x16, dim 256, 8 heads x 32, window 3x5x5), subsampled ----
    torch.manual_seed(42)
    ext = (1, 2, 2)
    m = Local3dAttention(ext, dim=256, heads=8, dim_head=32, use_checkpointing=False)
    sd = {k: seeded(100 + i, *v.shape) * 0.06 for i, (k, v) in enumerate(m.state_dict().items())}
    m.load_state_dict(sd)
    x = seeded(7, 2, 8, 16, 16, 256).requires_grad_(True)
    dout = seeded(8, 2, 8, 16, 16, 256)
    out = m(x, x)
    out.backward(dout)
    tok = slice(5, None, 16)     # every 16th token of the flattened (b s h w) axis
    np.savez_compressed(os.path.join(HERE, 'attn_c1.npz'), ext=np.array(ext), heads=8, dim_head=32,
                        x_sum=checksum(x.detach()), dout_sum=checksum(dout),
                        out=out.detach().reshape(-1, 256)[tok].numpy(),
                        dx=x.grad.reshape(-1, 256)[tok].numpy(),
                        **{'gradsum/' + k: checksum(v.grad) for k, v in m.named_parameters()},
                        **{'grad/' + k: v.grad.flatten()[::97].numpy() for k, v in m.named_parameters()})

    # ---- 3. config-4 widths on a reduced grid: attention core only -----------------
    ext = (2, 3, 3)
    heads, dh = 4, 128
    shape = (1, 6, 10, 10, heads * dh)
    qq = (seeded(11, *shape) * 0.5).requires_grad_(True)
    kk = (seeded(12, *shape) * 0.5).requires_grad_(True)
    vv = seeded(13, *shape).requires_grad_(True)
    dout = seeded(14, *shape)
    m = Local3dAttention(ext, dim=heads * dh, heads=heads, dim_head=dh, use_checkpointing=False)
    core = m.local_attention(kk, vv, qq)                       # [(bshw), heads, 1, d]
    core = core.reshape(*shape[:4], heads * dh)
    core.backward(dout)
    np.savez_compressed(os.path.join(HERE, 'core_c4_reduced.npz'), ext=np.array(ext), heads=heads, dim_head=dh,
                        shape=np.array(shape), q_sum=checksum(qq.detach()), out=core.detach().numpy()[0, ::2, ::3, ::3],
                        dq=qq.grad.numpy()[0, ::2, ::3, ::3], dk=kk.grad.numpy()[0, ::2, ::3, ::3],
                        dv=vv.grad.numpy()[0, ::2, ::3, ::3])

    # ---- 4. small transformer + denoiser head, fwd and grads -----------------------
    torch.manual_seed(3)
    cfg = dict(data_shape=(4, 6, 6), dim=32, num_classes=18, extents=(1, 1, 2), depth=2, heads=2, dim_head=16,
               mlp_dim=48)
    tr = Local3dAttentionTransformer(**cfg)
    head = torch.nn.Linear(32, 17)
    tokens = torch.randint(0, 18, (2, 4, 6, 6))
    target = torch.randint(0, 17, (2, 6, 6))
    feats = tr(tokens)
    logits = head(feats[:, -1])
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, 17), target.reshape(-1))
    loss.backward()
    grads = {'grad/transformer.' + k: v.grad.numpy() for k, v in tr.named_parameters()}
    grads.update({'grad/logit_proj.' + k: v.grad.numpy() for k, v in head.named_parameters()})
    sd = {'sd/transformer.' + k: v.detach().numpy() for k, v in tr.state_dict().items()}
    sd.update({'sd/logit_proj.' + k: v.detach().numpy() for k, v in head.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, 'denoiser_small.npz'), tokens=tokens.numpy(), target=target.numpy(),
                        feats=feats.detach().numpy(), logits=logits.detach().numpy(), loss=loss.item(),
                        cfg=np.array([4, 6, 6, 32, 17, 1, 1, 2, 2, 2, 16, 48]), **sd, **grads)

    # ---- 5. VQ: eval + training forward, single latent, K=512, D=64 (config 2 widths) ----
    torch.manual_seed(4)
    vq = VectorQuantizerEMA(64, 512)
    x = torch.randn(2, 16, 16, 64) * 0.9
    emb0 = vq.embedding.clone().numpy()
    vq.eval()
    qz, enc, loss_e, ppl_e = vq(x)
    eval_out = dict(idx_eval=enc.argmax(-1).numpy(), q_eval=qz.numpy(), loss_eval=loss_e.item(), ppl_eval=ppl_e.item(),
                    acc_err_eval=vq.accumulated_error.clone().numpy(), enc_sum=enc.sum().item(),
                    encode=vq.encode(x).numpy(), dist=vq.codebook_distance(x)[::37].numpy())
    vq.reset_stats()
    vq.train()
    qz, enc, loss_t, ppl_t = vq(x)
    np.savez_compressed(os.path.join(HERE, 'vq_c2.npz'), x=x.numpy(), embedding0=emb0, **eval_out,
                        loss_train=loss_t.item(), ppl_train=ppl_t.item(), embedding1=vq.embedding.numpy(),
                        cluster_size1=vq.cluster_size.numpy(), activation_count1=vq.activation_count.numpy(),
                        acc_err1=vq.accumulated_error.numpy())

    # ---- 5b. tie-break: duplicated and near-duplicated codes -----------------------
    torch.manual_seed(5)
    vq = VectorQuantizerEMA(16, 40)
    e = vq.embedding
    e[0, 7] = e[0, 3]
    e[0, 30] = e[0, 3]
    e[0, 21] = e[0, 20]
    e[0, 11] = e[0, 10] + 1e-3
    x = torch.cat([e[0, [3, 20, 10, 11, 39, 0]] + 0.01 * torch.randn(6, 16), torch.randn(250, 16)])
    vq.eval()
    np.savez_compressed(os.path.join(HERE, 'vq_ties.npz'), x=x.numpy(), embedding=e.numpy(),
                        encode=vq.encode(x).numpy())

    # ---- 5c. multi-latent codebooks (L=3) ------------------------------------------
    torch.manual_seed(6)
    vq = VectorQuantizerEMA(8, 12, num_latents=3)
    x = torch.randn(5, 7, 3, 8)
    emb0 = vq.embedding.clone().numpy()
    vq.train()
    qz, enc, loss_t, ppl_t = vq(x)
    idx = enc.argmax(-1)
    np.savez_compressed(os.path.join(HERE, 'vq_multilatent.npz'), x=x.numpy(), embedding0=emb0, idx=idx.numpy(),
                        q=qz.numpy(), loss=loss_t.item(), ppl=ppl_t.item(), embedding1=vq.embedding.numpy(),
                        cluster_size1=vq.cluster_size.numpy(), decode=vq.decode(idx.view(5, 7, 3)).numpy())

    # ---- 5d. 65 536 seeded vectors: indices only ------------------------------------
    vq = VectorQuantizerEMA(64, 512)
    vq.embedding.copy_(seeded(21, 1, 512, 64))
    x = seeded(22, 65536, 64)
    idx = torch.cat([vq.encode(x[i:i + 4096]) for i in range(0, 65536, 4096)])
    np.savez_compressed(os.path.join(HERE, 'vq_64k.npz'), x_sum=checksum(x), idx=idx.numpy().astype(np.int16))

    # ---- 5e. config 2: 64x64 frames -> conv encoder (downscale_steps=2) -> 16x16x64 latents ----
    torch.manual_seed(7)
    enc_net = SimpleResidualEncoder(1, 64, 2, 128)
    frames = torch.rand(2, 1, 64, 64)
    with torch.no_grad():
        lat = enc_net(frames).permute(0, 2, 3, 1).contiguous()
    vq = VectorQuantizerEMA(64, 512)
    vq.embedding.copy_(seeded(23, 1, 512, 64) * lat.std() + lat.mean())
    np.savez_compressed(os.path.join(HERE, 'vq_c2_latents.npz'), latents=lat.numpy(), embedding=vq.embedding.numpy(),
                        encode=vq.encode(lat).view(2, 16, 16).numpy())
    # ---- 6. VqAutoEncoder (train_vqae.py:22-55 cannot be imported: matplotlib): the reference encoder, quantizer and
    #         decoder composed under the reference's attribute names; encode / decode / forward in eval and train mode ----
    class RefVqAutoEncoder(torch.nn.Module):
        def __init__(self, embedding_dim, num_embeddings, downscale_steps, hidden_planes, in_channels):
            super().__init__()
            self.encoder = SimpleResidualEncoder(in_channels, embedding_dim, downscale_steps, hidden_planes)
            self.decoder = SimpleResidualDecoder([hidden_planes] * downscale_steps, in_channels=embedding_dim,
                                                 out_channels=in_channels)
            self.vq = VectorQuantizerEMA(embedding_dim, num_embeddings)

    torch.manual_seed(8)
    ae = RefVqAutoEncoder(16, 64, 2, 32, 1)
    for m_ in ae.modules():                         # non-trivial BatchNorm statistics
        if isinstance(m_, torch.nn.BatchNorm2d):
            m_.running_mean.normal_(0, 0.1)
            m_.running_var.uniform_(0.5, 1.5)
            m_.weight.data.uniform_(0.5, 1.5)
            m_.bias.data.normal_(0, 0.1)
    frames = torch.rand(3, 1, 32, 32)
    sd0 = {'sd/' + k: v.detach().clone().numpy() for k, v in ae.state_dict().items()}
    ae.eval()
    with torch.no_grad():
        h = ae.encoder(frames).permute(0, 2, 3, 1)
        ae.vq.embedding.copy_(torch.randn(1, 64, 16) * h.std() + h.mean())
        sd0['sd/vq.embedding'] = ae.vq.embedding.clone().numpy()
        z = ae.vq.encode(h).view(h.shape[:-1])                                      # train_vqae.py:45-49
        dec = ae.decoder(ae.vq.decode(z).permute(0, 3, 1, 2))                       # :51-55
        q, _, latent_loss, ppl = ae.vq.forward(h)                                   # :33-43
        recon = ae.decoder(q.permute(0, 3, 1, 2).contiguous())
    ae.train()                                      # main.py never calls .eval() on the tokenizer: batch statistics
    with torch.no_grad():
        z_train = ae.vq.encode(ae.encoder(frames).permute(0, 2, 3, 1)).view(3, 8, 8)
    np.savez_compressed(os.path.join(HERE, 'vqae_small.npz'), frames=frames.numpy(), latents=h.numpy(), z=z.numpy(),
                        decoded=dec.numpy(), recon=recon.numpy(), latent_loss=latent_loss.item(), ppl=ppl.item(),
                        z_train=z_train.numpy(), cfg=np.array([16, 64, 2, 32, 1]), **sd0)

    # ---- 7. sparse-context denoiser (minecraft/sparse_diffusion.py:75-111 cannot be imported: minerl): the reference's
    #         dense Transformer (minecraft/transformer.py) under the wrapper's attribute names ----
    sys.path.insert(0, '/root/reference/minecraft')
    import importlib
    ref_tr = importlib.import_module('transformer')
    sys.path.pop(0)

    class RefSparse(torch.nn.Module):
        def __init__(self, shape, dim, num_classes, depth, dim_head, mlp_dim, heads):
            super().__init__()
            S, H, W = shape
            self.pos_emb_s = torch.nn.Embedding(S, dim)
            self.pos_emb_h = torch.nn.Embedding(H, dim)
            self.pos_emb_w = torch.nn.Embedding(W, dim)
            self.embedding = torch.nn.Embedding(num_classes + 1, dim)
            self.transformer = ref_tr.Transformer(dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim)
            self.logit_proj = torch.nn.Linear(dim, num_classes)

    torch.manual_seed(9)
    shape, dim, K, depth, dh, mlp, heads = (4, 6, 6), 32, 20, 2, 16, 48, 2
    sp = RefSparse(shape, dim, K, depth, dh, mlp, heads)
    idx = torch.stack([torch.randperm(4 * 6 * 6)[:24] for _ in range(3)])
    tok = torch.randint(0, K + 1, (3, 24))
    tgt = torch.randint(0, K, (3, 24))
    W_, H_ = shape[2], shape[1]
    pos = sp.pos_emb_s(idx.div(H_ * W_, rounding_mode='trunc')) + sp.pos_emb_h(idx.div(W_, rounding_mode='trunc') % H_) \
        + sp.pos_emb_w(idx % W_)                                                     # sparse_diffusion.py:100-105
    logits = sp.logit_proj(sp.transformer(sp.embedding(tok) + pos))                  # :107-111
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, K), tgt.reshape(-1))
    loss.backward()
    np.savez_compressed(os.path.join(HERE, 'sparse_small.npz'), tokens=tok.numpy(), indices=idx.numpy(), target=tgt.numpy(),
                        logits=logits.detach().numpy(), loss=loss.item(), cfg=np.array([*shape, dim, K, depth, dh, mlp, heads]),
                        **sd_np(sp), **{'grad/' + k: v.grad.numpy() for k, v in sp.named_parameters()})

    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
