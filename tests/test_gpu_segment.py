"""Segment reductions (`-m gpu`): the plain kernels (one CSR, and two CSRs of one edge tensor in one launch) against
(a) the ReLU-masked variant fed an all-ones mask -- same summation order, so the comparison is bit-exact -- and (b) an
fp64 index_add (my_gnn_layer.py:301-307 `aggr='mean'` / the scatter of its backward), on full, edge-dropped, wide,
tiny and strided cases."""
import numpy as np
import pytest
import torch

from relpose_gnn_b200 import graph as G, ops

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def dev():
    return torch.device("cuda:0")


def make_graph(N, Gn, keep_frac, seed):
    src, dst = G.fc_template(N)
    if keep_frac < 1.0:
        H = N * (N - 1) // 2
        rng = np.random.RandomState(seed)
        keep = rng.rand(H) < keep_frac
        keep[0] = True
        return G.GraphBatch.fully_connected(Gn, N, dev(), keep)
    return G.GraphBatch(src, dst, Gn, N, dev())


def node_of(graph, which):
    ei = graph.edge_index()
    s, d = ei[0], ei[1]
    return {"in": d, "out": s, "min": torch.minimum(s, d), "max": torch.maximum(s, d)}[which]


def ref_sum(v, graph, which, Nt):
    out = torch.zeros(Nt, v.size(1), dtype=torch.float64, device=v.device)
    out.index_add_(0, node_of(graph, which), v.double())
    return out


CASES = [
    # N, G, keep fraction, D
    (9, 300, 1.0, 512),        # full template
    (9, 301, 0.5, 512),        # edge dropout: short ragged CSR rows, isolated nodes possible
    (17, 70, 1.0, 512),        # 272 edge rows per graph
    (17, 75, 0.5, 512),
    (9, 64, 0.5, 64),          # narrow features: several node rows per thread group
    (3, 1000, 1.0, 128),
    (2, 5, 1.0, 512),          # fewer graphs than CTAs
    (33, 9, 1.0, 256),         # 1056 edge rows per graph
]


@pytest.mark.parametrize("N,Gn,kf,D", CASES)
def test_segment_sum_is_bit_identical_to_the_masked_variant_and_matches_fp64(N, Gn, kf, D):
    graph = make_graph(N, Gn, kf, 11 * N + Gn)
    Et, Nt = graph.n_edge_rows, graph.n_node_rows
    gen = torch.Generator().manual_seed(N * 1000 + Gn)
    v = torch.randn(Et, D, generator=gen).to(dev()).bfloat16()
    ones = torch.ones_like(v)
    for which in ("in", "out", "min", "max"):
        got = torch.full((Nt, D), 7.0, dtype=BF16, device=dev())
        want = torch.full((Nt, D), -7.0, dtype=BF16, device=dev())
        ops.segment_sum(v, graph, which, got)
        ops.segment_sum(v, graph, which, want, mask=ones)           # masked variant of the kernel
        assert torch.equal(got, want), which
        ref = ref_sum(v, graph, which, Nt)
        err = (got.double() - ref).abs().max().item()
        assert err <= 2 ** -8 * max(ref.abs().max().item(), 1.0) + 1e-6, (which, err)   # one bf16 rounding
    # mean aggregation: the 1 / in-degree scale (0 for isolated nodes) through the same kernel
    got = torch.empty(Nt, D, dtype=BF16, device=dev())
    ops.aggregate_mean(v, graph, got)
    deg = torch.bincount(node_of(graph, "in"), minlength=Nt).double().unsqueeze(1)
    ref = ref_sum(v, graph, "in", Nt) / deg.clamp(min=1)
    assert (got.double() - ref).abs().max().item() <= 2 ** -8 * max(ref.abs().max().item(), 1.0) + 1e-6
    assert float(got.float()[(deg == 0).squeeze(1)].abs().sum()) == 0.0


@pytest.mark.parametrize("N,Gn,kf,D", CASES)
def test_two_csr_pass_matches_two_single_passes(N, Gn, kf, D):
    graph = make_graph(N, Gn, kf, 5 * N + Gn)
    Et, Nt = graph.n_edge_rows, graph.n_node_rows
    gen = torch.Generator().manual_seed(N * 77 + Gn)
    v = torch.randn(Et, D, generator=gen).to(dev()).bfloat16()
    ones = torch.ones_like(v)
    for wa, wb in (("out", "in"), ("min", "max")):
        both = torch.zeros(Nt, 2 * D, dtype=BF16, device=dev())      # strided outputs: two halves of one buffer
        ops.segment_sum2(v, graph, wa, both[:, :D], wb, both[:, D:])
        a = torch.empty(Nt, D, dtype=BF16, device=dev())
        b = torch.empty(Nt, D, dtype=BF16, device=dev())
        ops.segment_sum(v, graph, wa, a, mask=ones)
        ops.segment_sum(v, graph, wb, b, mask=ones)
        assert torch.equal(both[:, :D], a) and torch.equal(both[:, D:], b), (wa, wb)


def test_segment_sum_on_a_strided_input_view():
    """Input rows that are a column window of a wider tensor (ldv != D)."""
    N, Gn, D = 9, 123, 256
    graph = make_graph(N, Gn, 0.5, 3)
    Et, Nt = graph.n_edge_rows, graph.n_node_rows
    wide = torch.randn(Et, 3 * D, generator=torch.Generator().manual_seed(9)).to(dev()).bfloat16()
    v = wide[:, D:2 * D]
    got = torch.empty(Nt, D, dtype=BF16, device=dev())
    want = torch.empty(Nt, D, dtype=BF16, device=dev())
    ops.segment_sum(v, graph, "in", got)
    ops.segment_sum(v.contiguous(), graph, "in", want, mask=torch.ones(Et, D, dtype=BF16, device=dev()))
    assert torch.equal(got, want)


def test_segment_sum2_is_deterministic_at_size():
    """Config C size (4096 x 9, edge dropout): two runs agree bit for bit and with the fp64 sum."""
    graph = make_graph(9, 4096, 0.5, 1)
    Et, Nt, D = graph.n_edge_rows, graph.n_node_rows, 512
    v = torch.randn(Et, D, generator=torch.Generator().manual_seed(4)).to(dev()).bfloat16()
    outs = []
    for _ in range(2):
        o = torch.empty(Nt, 2 * D, dtype=BF16, device=dev())
        ops.segment_sum2(v, graph, "min", o[:, :D], "max", o[:, D:])
        outs.append(o)
    assert torch.equal(outs[0], outs[1])
    ref = torch.cat([ref_sum(v, graph, "min", Nt), ref_sum(v, graph, "max", Nt)], 1)
    assert (outs[0].double() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-6
