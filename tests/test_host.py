"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header declares,
the module mirrors keep the reference's state_dict, the product refuses CPU tensors, the product never imports
the oracle, and the data-parallel gradient buckets are equivalent to a single-process global batch (gloo, world 2)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "vitgan_b200.h")).read()
    declared = set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(os.path.join(ROOT, "vit-gan_b200", "libvitgan_b200.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vitgan_b200.h but not exported"
    from vitgan_b200 import lib as L
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert L.lib.vg_version() == 5 and L.lib.vg_last_error() is not None


def test_ctypes_signatures_match_header_prototypes():
    """Every ctypes binding has exactly the parameter list of its prototype in include/vitgan_b200.h (count and kind:
    pointer / 32-bit int / 64-bit int / float).  A missing trailing pointer is passed by ctypes as a C int, i.e. truncated to
    32 bits -- the stream handle of vg_sln_bwd was lost that way on non-default streams."""
    import ctypes as C
    from vitgan_b200 import lib as L
    hdr = open(os.path.join(ROOT, "include", "vitgan_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = dict(re.findall(r"\b(?:int|const char\*)\s+(vg_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S))
    assert set(protos) == set(L.SIGNATURES)

    def kind(param):
        p = " ".join(param.split())
        if "*" in p:
            return "ptr"
        base = p.rsplit(" ", 1)[0] if " " in p else p
        return {"int": "i32", "int64_t": "i64", "float": "f32", "unsigned": "i32"}[base.replace("const ", "")]

    ck = {C.c_int: "i32", C.c_int64: "i64", C.c_float: "f32", C.c_void_p: "ptr", C.c_char_p: "ptr"}
    for name, params in protos.items():
        want = [] if params.strip() in ("", "void") else [kind(x) for x in params.split(",")]
        got = [ck.get(t, "ptr") for t in L.SIGNATURES[name]]
        assert got == want, (name, got, want)


def test_struct_layout_matches_header():
    """sizeof(vg_gemm_args) as compiled == the ctypes mirror (guards against silent ABI drift)."""
    import ctypes
    from vitgan_b200 import lib as L
    src = '#include "vitgan_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu", sizeof(vg_gemm_args), offsetof(vg_gemm_args, accumulate));return 0;}'
    exe = "/tmp/vg_abi_probe"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    size, off = map(int, subprocess.check_output([exe]).split())
    assert size == ctypes.sizeof(L.GemmArgs) and off == L.GemmArgs.accumulate.offset


def test_mirrors_keep_reference_state_dict_and_init():
    import vitgan_b200 as vb
    from oracle import harness, v1 as o1, v2 as o2
    torch.manual_seed(0)
    gan = vb.v2.ViTGAN(vb.v2.Config(batch_size=3072))
    ref = o2.init_vitgan(o2.V2Config(batch_size=3072), seed=0)      # == reference init (tests/test_oracle_vs_reference.py)
    assert list(gan.state_dict()) and all(torch.equal(ref[k], v) for k, v in gan.state_dict().items()) and len(ref) == len(gan.state_dict())
    torch.manual_seed(0)
    G, D = vb.v1.Generator(vb.v1.V1Config(image_size=32)), vb.v1.Discriminator(vb.v1.V1Config(image_size=32))
    orc = harness.OracleV1(o1.V1Config(image_size=32), seed=0)
    assert all(torch.equal(orc.p["generator." + k], v) for k, v in G.state_dict().items())
    assert all(torch.equal(orc.p["discriminator." + k], v) for k, v in D.state_dict().items())
    assert len(G.state_dict()) + len(D.state_dict()) == len(orc.p)


def test_no_cpu_fallback_and_no_oracle_import():
    import vitgan_b200 as vb
    gan = vb.v2.ViTGAN(vb.v2.Config(embeddings_dimension=32, transformer_blocks_count=1, image_size=16, batch_size=768))
    # reference default p = 0.1 in training mode (src/v2/modules.py:99,179-180): unsupported => error, never silently skipped
    with pytest.raises(RuntimeError, match="nn.Dropout"):
        gan.discriminator(torch.randn(2, 3, 16, 16))
    gan.eval()
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        gan.discriminator(torch.randn(2, 3, 16, 16))
    out = subprocess.check_output([sys.executable, "-c",
                                   "import sys; sys.path.insert(0, %r); import vitgan_b200; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))" % ROOT])
    assert out.strip() == b"False"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vit-gan_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(dirpath, f)).read(), re.M), f


def test_patch_binds_reference_instances():
    from oracle import refimport
    if not refimport.available():
        pytest.skip("/root/reference not present")
    import vitgan_b200 as vb
    gan, c = refimport.build_v2(seed=0)
    keys = list(gan.state_dict())
    params = [id(p) for p in gan.parameters()]
    n = vb.patch.patch_v2(gan)
    assert n == 2 * (1 + 6 * 2 + 1 + 1) + 2                       # per ViT: embed, 6x(encoder+attention), classifier, vit; + G, D
    assert list(gan.state_dict()) == keys and [id(p) for p in gan.parameters()] == params
    with pytest.raises(RuntimeError, match="CUDA tensor"):        # patched forward is the CUDA one: no silent CPU path
        gan.discriminator(torch.randn(2, 3, 32, 32))
    G, D = refimport.build_v1(32, seed=0)
    assert vb.patch.patch_v1(G, D) > 20
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        D(torch.randn(2, 3, 32, 32))


def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    import vitgan_b200 as vb
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))
    unused = torch.nn.Linear(3, 3)            # parameters whose hooks never fire must still be reduced by finish()
    net.add_module("unused", unused)
    flat = vb.train.FlatNet(net)
    buckets = vb.train.GradBuckets(flat, n_buckets=2)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 12, generator=g), torch.randn(8, 1, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    flat.zero_grad()
    # pass 1 (not armed: accumulates locally, like D's first backward), pass 2 armed
    torch.nn.functional.mse_loss(net[:4](xs), ys).backward()
    buckets.arm()
    torch.nn.functional.mse_loss(net[:4](xs * 0.5), ys).backward()
    buckets.finish()
    ret[rank] = flat.flat_grad.clone()
    dist.destroy_process_group()


def test_grad_buckets_gloo_world2_equals_global_batch():
    import torch.multiprocessing as mp
    import vitgan_b200 as vb
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    # single-process reference on the global batch (mean-of-means with equal shards == global mean)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))
    net.add_module("unused", torch.nn.Linear(3, 3))
    flat = vb.train.FlatNet(net)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 12, generator=g), torch.randn(8, 1, generator=g)
    flat.zero_grad()
    # NOTE pass 1 is NOT all-reduced in the runner (local accumulation) but IS summed into the same flat buffer,
    # so after the armed all-reduce both passes are averaged: identical to the global-batch gradient of both losses.
    torch.nn.functional.mse_loss(net[:4](x), y).backward()
    torch.nn.functional.mse_loss(net[:4](x * 0.5), y).backward()
    assert torch.allclose(ret[0], ret[1])
    assert torch.allclose(ret[0], flat.flat_grad, atol=1e-6), (ret[0] - flat.flat_grad).abs().max()


def test_merged_discriminator_pass_is_the_same_step_host_logic():
    """train.gan_step(merge_d_passes=True) -- D(real) and D(fake.detach()) as ONE pass over the concatenated batch with
    loss = mean_real + mean_fake -- is the reference's two-pass sequence (src/v2/training.py:177-197) up to fp32 summation
    order: same three losses and same parameters after two optimizer steps.  Host-side schedule only (plain torch modules on
    the CPU stand in for the kernels; the GPU tests pin the real thing against the reference-generated loss curves)."""
    import copy
    import vitgan_b200 as vb
    torch.manual_seed(0)
    g0 = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 12), torch.nn.Unflatten(1, (3, 2, 2)))
    d0 = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.Linear(16, 5))
    data = [(torch.randn(6, 3, 2, 2), torch.randn(6, 3, 2, 2)) for _ in range(2)]
    out = {}
    for merged in (False, True):
        g, d = copy.deepcopy(g0).double(), copy.deepcopy(d0).double()
        go, do = torch.optim.AdamW(g.parameters(), lr=1e-2), torch.optim.AdamW(d.parameters(), lr=1e-2)
        losses = [torch.stack([t.reshape(()) for t in vb.train.gan_step(g, d, go, do, r.double(), n.double(), "ce", merge_d_passes=merged)])
                  for r, n in data]
        out[merged] = (torch.stack(losses), torch.cat([p.detach().reshape(-1) for p in list(g.parameters()) + list(d.parameters())]))
    # gan_step evaluates the loss head in fp32 (`.float()` on the logits), so fp32 rounding is the floor
    assert torch.allclose(out[True][0], out[False][0], rtol=1e-5, atol=1e-6)
    assert torch.allclose(out[True][1], out[False][1], rtol=1e-4, atol=1e-5)
