"""The chain rule the folded train step relies on (csrc/train.cu `unfold_jobs`): gradients of the folded
weights W23 = W2 W3, a23 = A2 a3, UH = U2 U3 H1, HO = H2 H3 (and their biases) pushed back onto the flax parameters.
Pure float64 algebra against torch autograd -- it pins the formulas the CUDA job table encodes term by term; the CUDA
step itself is compared with the float64 oracle in tests/test_gpu_train.py."""
import torch


def _rand(*shape, g):
    return torch.randn(*shape, generator=g, dtype=torch.float64)


def test_unfolded_gradients_equal_autograd_through_the_products():
    g = torch.Generator().manual_seed(7)
    no = 2
    names = ["W2", "b2", "W3", "b3", "A2", "ba2", "a3", "ba3", "U2", "bu2", "U3", "bu3", "H1", "bh1", "H2", "bh2", "H3", "bh3"]
    shapes = [(256, 256), (256,), (256, 128), (128,), (128, 128), (128,), (128, 1), (1,), (256, 256), (256,), (256, 128),
              (128,), (128, 256), (256,), (256, 256), (256,), (256, no), (no,)]
    P = {n: (_rand(*s, g=g) * 0.1).requires_grad_(True) for n, s in zip(names, shapes)}

    def fold(P):
        W23 = P["W2"] @ P["W3"]
        b23 = P["b2"] @ P["W3"] + P["b3"]
        a23 = (P["A2"] @ P["a3"])[:, 0]
        c23 = P["ba2"] @ P["a3"][:, 0] + P["ba3"][0]
        Q = P["U2"] @ P["U3"]
        bq = P["bu2"] @ P["U3"] + P["bu3"]
        UH = Q @ P["H1"]
        buh = bq @ P["H1"] + P["bh1"]
        HO = P["H2"] @ P["H3"]
        bho = P["bh2"] @ P["H3"] + P["bh3"]
        return dict(W23=W23, b23=b23, a23=a23, c23=c23, UH=UH, buh=buh, HO=HO, bho=bho, Q=Q, bq=bq)

    F = fold(P)
    keys = ["W23", "b23", "a23", "c23", "UH", "buh", "HO", "bho"]
    Gf = {k: _rand(*F[k].shape, g=g) if F[k].dim() else _rand(1, g=g)[0] for k in keys}   # any upstream gradient
    scalar = sum((F[k] * Gf[k]).sum() for k in keys)
    auto = dict(zip(names, torch.autograd.grad(scalar, [P[n] for n in names])))

    with torch.no_grad():
        D = {n: P[n].detach() for n in names}
        Q, bq = F["Q"].detach(), F["bq"].detach()
        T = Gf["UH"] @ D["H1"].T                       # [256, 128]
        t = Gf["buh"] @ D["H1"].T                      # [128]
        mine = {
            "W2": Gf["W23"] @ D["W3"].T,
            "W3": D["W2"].T @ Gf["W23"] + torch.outer(D["b2"], Gf["b23"]),
            "b2": Gf["b23"] @ D["W3"].T,
            "b3": Gf["b23"],
            "A2": torch.outer(Gf["a23"], D["a3"][:, 0]),
            "a3": (D["A2"].T @ Gf["a23"] + D["ba2"] * Gf["c23"])[:, None],
            "ba2": D["a3"][:, 0] * Gf["c23"],
            "ba3": Gf["c23"].reshape(1),
            "U2": T @ D["U3"].T,
            "U3": D["U2"].T @ T + torch.outer(D["bu2"], t),
            "H1": Q.T @ Gf["UH"] + torch.outer(bq, Gf["buh"]),
            "bu2": t @ D["U3"].T,
            "bu3": t,
            "bh1": Gf["buh"],
            "H2": Gf["HO"] @ D["H3"].T,
            "H3": D["H2"].T @ Gf["HO"] + torch.outer(D["bh2"], Gf["bho"]),
            "bh2": Gf["bho"] @ D["H3"].T,
            "bh3": Gf["bho"],
        }
    for n in names:
        torch.testing.assert_close(mine[n], auto[n], rtol=1e-10, atol=1e-12, msg=n)


def test_folded_forward_is_the_layer_by_layer_forward():
    """relu(x) W2 W3-style blocks: folding the activation-free tail changes rounding only."""
    g = torch.Generator().manual_seed(3)
    x1 = torch.relu(_rand(37, 256, g=g))
    W2, b2, W3, b3 = _rand(256, 256, g=g) * 0.1, _rand(256, g=g), _rand(256, 128, g=g) * 0.1, _rand(128, g=g)
    seq = (x1 @ W2 + b2) @ W3 + b3
    fold = x1 @ (W2 @ W3) + (b2 @ W3 + b3)
    torch.testing.assert_close(fold, seq, rtol=1e-12, atol=1e-12)
