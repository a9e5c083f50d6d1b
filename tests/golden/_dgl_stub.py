"""Pure-torch stand-in for the handful of DGL / jax symbols that the reference's
``code/nn_module.py`` and ``code/md_module.py`` touch.

TEST INFRASTRUCTURE ONLY.  It exists so that ``make_golden.py`` can import and
execute the *unmodified* reference modules from ``/root/reference`` on CPU in a
container that has neither DGL nor jax installed, in order to generate the golden
vectors committed under ``tests/golden/``.  Nothing in the product package
imports this file.

Semantics implemented (DGL 0.7 behaviour, see SURVEY.md section 8a note 1):
  * ``dgl.graph((src, dst))`` - directed multigraph container, edge order kept
  * ``g.edges()`` -> (src, dst)
  * ``g.edata / g.srcdata / g.dstdata / g.ndata`` - plain dicts
  * ``g.local_scope()`` - context manager restoring the dicts on exit
  * ``g.is_block`` - False
  * ``g.update_all(fn.src_mul_edge(u, e, m), fn.sum(m, out))`` - index_add_
  * ``g.add_self_loop()`` - returns a NEW graph (so the reference's bare call is a no-op)
  * ``dgl.batch`` - block-diagonal union with node-id offsets
  * ``dgl.add_reverse_edges`` and ``g.has_edges_between``
"""
import contextlib
import sys
import types

import torch


class _Msg:
    def __init__(self, kind, *names):
        self.kind = kind
        self.names = names


class DGLGraph:
    is_block = False

    def __init__(self, src, dst, num_nodes=None):
        self._src = torch.as_tensor(src).long()
        self._dst = torch.as_tensor(dst).long()
        if num_nodes is None:
            num_nodes = int(max(self._src.max().item(), self._dst.max().item())) + 1 if self._src.numel() else 0
        self._n = num_nodes
        self.edata = {}
        self.ndata = {}
        self.srcdata = self.ndata
        self.dstdata = self.ndata

    def edges(self):
        return self._src, self._dst

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_dst_nodes(self):
        return self._n

    @contextlib.contextmanager
    def local_scope(self):
        e, n = dict(self.edata), dict(self.ndata)
        try:
            yield
        finally:
            self.edata.clear(); self.edata.update(e)
            self.ndata.clear(); self.ndata.update(n)

    def update_all(self, msg, red):
        assert msg.kind == 'u_mul_e' and red.kind == 'sum'
        u, e, m = msg.names
        _, out = red.names
        h = self.ndata[u]
        mval = h[self._src] * self.edata[e]
        res = torch.zeros((self._n,) + tuple(mval.shape[1:]), dtype=mval.dtype)
        res.index_add_(0, self._dst, mval)
        self.ndata[out] = res

    def add_self_loop(self):
        ar = torch.arange(self._n)
        return DGLGraph(torch.cat([self._src, ar]), torch.cat([self._dst, ar]), self._n)

    def has_edges_between(self, u, v):
        key = set(zip(self._src.tolist(), self._dst.tolist()))
        u = torch.as_tensor(u).long().tolist()
        v = torch.as_tensor(v).long().tolist()
        return torch.tensor([(a, b) in key for a, b in zip(u, v)], dtype=torch.bool)


def graph(data, num_nodes=None):
    src, dst = data
    return DGLGraph(src, dst, num_nodes)


def batch(graphs):
    off = 0
    srcs, dsts, ed = [], [], {}
    for g in graphs:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        off += g._n
    out = DGLGraph(torch.cat(srcs), torch.cat(dsts), off)
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], dim=0)
    return out


def add_reverse_edges(g):
    return DGLGraph(torch.cat([g._src, g._dst]), torch.cat([g._dst, g._src]), g._n)


def install():
    """Register fake ``dgl`` / ``jax`` / ``jax_md`` modules in ``sys.modules``."""
    dgl = types.ModuleType('dgl')
    dgl.graph = graph
    dgl.batch = batch
    dgl.add_reverse_edges = add_reverse_edges
    dgl.DGLGraph = DGLGraph
    dgl_nn = types.ModuleType('dgl.nn')
    dgl_fn = types.ModuleType('dgl.function')
    dgl_fn.src_mul_edge = lambda u, e, m: _Msg('u_mul_e', u, e, m)
    dgl_fn.u_mul_e = dgl_fn.src_mul_edge
    dgl_fn.sum = lambda m, out: _Msg('sum', m, out)
    dgl_ops = types.ModuleType('dgl.ops')
    dgl_ops.edge_softmax = None
    dgl_utils = types.ModuleType('dgl.utils')
    dgl_utils.expand_as_pair = None
    dgl.nn, dgl.function, dgl.ops, dgl.utils = dgl_nn, dgl_fn, dgl_ops, dgl_utils
    sys.modules.update({'dgl': dgl, 'dgl.nn': dgl_nn, 'dgl.function': dgl_fn,
                        'dgl.ops': dgl_ops, 'dgl.utils': dgl_utils})

    # md_module.py does `import jax`, `from jax_md import space, partition`, uses @jax.jit
    # at import time; get_neighbor / pair_distance themselves are pure torch.
    jax = types.ModuleType('jax')
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    jax.vmap = lambda f, *a, **k: f
    jax.partial = lambda f, *a, **k: f
    jnp = types.ModuleType('jax.numpy')
    jnp.ndarray = object
    jax.numpy = jnp
    jax_md = types.ModuleType('jax_md')
    space = types.ModuleType('jax_md.space')
    space.pairwise_displacement = None
    partition = types.ModuleType('jax_md.partition')
    jax_md.space, jax_md.partition = space, partition
    sys.modules.update({'jax': jax, 'jax.numpy': jnp, 'jax_md': jax_md,
                        'jax_md.space': space, 'jax_md.partition': partition})
