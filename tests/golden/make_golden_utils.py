"""Generate tests/golden/utils_*.npz by RUNNING THE REFERENCE's sort_edge_index / coalesce
(/root/reference/gammagl/utils/sort_edge_index.py, coalesce.py), in the authoring container only:

    python tests/golden/make_golden_utils.py

The two functions import `tensorlayerx` (absent here) for six thin tensor helpers; their source is executed unmodified
apart from the import lines, inside a namespace that supplies those helpers on torch (what tensorlayerx's torch backend
does) and the reference's own pure-torch `mpops.segment_sum` (gammagl/mpops/torch.py loaded by path).  Contents:

  utils_kat.npz        the literal cases of /root/reference/tests/utils/test_sort_edge_index.py:7-16 and
                       tests/utils/test_coalesce.py:5-21 (inputs + expectations; asserted against the reference here)
  utils_rand_*.npz     seeded random edge lists: one without duplicate edges (attributes comparable element-wise) and
                       two with duplicates (integer-valued float attributes, so the merged sums do not depend on the order
                       in which the reference's unstable argsort leaves equal keys)
"""
import importlib.util
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/gammagl'


def _load_mpops():
    spec = importlib.util.spec_from_file_location('ref_mpops_torch', os.path.join(REF, 'mpops', 'torch.py'))
    m = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(m)
    except ImportError:
        pass
    return m


def _tlx_shim():
    tlx = types.SimpleNamespace()
    tlx.int64 = torch.int64
    tlx.is_tensor = lambda x: isinstance(x, torch.Tensor)
    tlx.convert_to_numpy = lambda x: x.detach().cpu().numpy()
    tlx.convert_to_tensor = lambda x, dtype=None: torch.as_tensor(x, dtype=dtype)
    tlx.reduce_max = lambda x: x.max()

    def gather(x, indices, axis=0):
        return torch.index_select(x, axis, torch.as_tensor(indices, dtype=torch.int64))
    tlx.gather = gather
    tlx.concat = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
    tlx.stack = lambda xs, axis=0: torch.stack(list(xs), dim=axis)
    tlx.all = torch.all
    tlx.ops = types.SimpleNamespace(is_tensor=tlx.is_tensor, argsort=lambda x, descending=False: torch.argsort(x, descending=descending),
                                    convert_to_numpy=tlx.convert_to_numpy)
    return tlx


def _exec_reference(relpath, ns):
    src = open(os.path.join(REF, relpath)).read()
    body = '\n'.join(line for line in src.splitlines() if not line.startswith(('import ', 'from ')))
    exec(compile(body, os.path.join(REF, relpath), 'exec'), ns)


def load_reference_utils():
    ns = {'np': np, 'tlx': _tlx_shim(), 'mpops': _load_mpops(), 'copy': __import__('copy').copy}
    _exec_reference('utils/check.py', ns)
    _exec_reference('utils/num_nodes.py', ns)
    _exec_reference('utils/sort_edge_index.py', ns)
    _exec_reference('utils/coalesce.py', ns)
    _exec_reference('utils/undirected.py', ns)
    return ns['sort_edge_index'], ns['coalesce'], ns['to_undirected'], ns['is_undirected']


def main():
    sort_edge_index, coalesce, to_undirected, is_undirected = load_reference_utils()
    # ---- KATs -------------------------------------------------------------------------------------
    ei = torch.tensor([[2, 1, 1, 0], [1, 2, 0, 1]])
    attr = torch.tensor([[1], [2], [3], [4]])
    out = sort_edge_index(ei)
    assert out.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]                       # test_sort_edge_index.py:10-11
    out2 = sort_edge_index(ei, attr)
    assert out2[0].tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]] and out2[1].tolist() == [[4], [3], [2], [1]]   # :13-15
    cei = torch.tensor([[0, 1, 0, 1], [1, 0, 1, 0]])
    cattr = torch.tensor([[1, 0], [2, 3], [-5, 4], [0, 2]], dtype=torch.float32)
    c1 = coalesce(cei)
    assert c1.tolist() == [[0, 1], [1, 0]]                                   # test_coalesce.py:13-14
    c2 = coalesce(cei, cattr)
    assert c2[0].tolist() == [[0, 1], [1, 0]] and c2[1].tolist() == [[-4.0, 4.0], [2.0, 5.0]]              # :18-21
    # tests/utils/test_undirected.py:5-36
    r1, c1 = torch.tensor([0, 1, 0]), torch.tensor([1, 0, 0])
    assert is_undirected(torch.stack([r1, c1])) and is_undirected(torch.stack([r1, c1]), torch.tensor([0, 0, 1]))
    assert not is_undirected(torch.stack([r1, c1]), torch.tensor([0, 1, 1]))
    r2, c2 = torch.tensor([0, 1, 1]), torch.tensor([1, 0, 2])
    assert not is_undirected(torch.stack([r2, c2]))
    uattr = torch.tensor([[1, 0], [2, 3], [-5, 4]], dtype=torch.float32)
    uei, ua = to_undirected(edge_index=torch.stack([r2, c2]), edge_attr=uattr, reduce='add')
    assert uei.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]] and ua.tolist() == [[3., 3.], [3., 3.], [-5., 4.], [-5., 4.]]
    np.savez(os.path.join(HERE, 'utils_kat.npz'), und_ei=torch.stack([r2, c2]).numpy(), und_attr=uattr.numpy(), und_out=uei.numpy(),
             und_out_attr=ua.numpy(), sort_ei=ei.numpy(), sort_attr=attr.numpy(), sort_out=out.numpy(),
             sort_out_attr=out2[1].numpy(), co_ei=cei.numpy(), co_attr=cattr.numpy(), co_out=c2[0].numpy(),
             co_out_attr=c2[1].numpy())
    # ---- randomized -------------------------------------------------------------------------------
    for name, N, E, dup, seed in (('nodup', 400, 8000, False, 1), ('dup', 150, 12000, True, 2), ('dup_small', 17, 2000, True, 3)):
        g = torch.Generator().manual_seed(seed)
        if dup:
            e = torch.randint(0, N, (2, E), generator=g)
        else:
            flat = torch.randperm(N * N, generator=g)[:E]
            e = torch.stack([flat // N, flat % N])
        a = torch.randint(-8, 9, (E, 3), generator=g).float()
        rec = {'ei': e.numpy(), 'attr': a.numpy(), 'N': np.int64(N)}
        for sbr in (True, False):
            s_ei, s_at = sort_edge_index(e.clone(), a.clone(), N, sort_by_row=sbr)
            c_ei, c_at = coalesce(e.clone(), a.clone(), N, sort_by_row=sbr)
            tag = 'row' if sbr else 'col'
            rec['sort_%s_ei' % tag] = s_ei.numpy()
            if not dup:
                rec['sort_%s_attr' % tag] = s_at.numpy()
            rec['co_%s_ei' % tag] = c_ei.numpy()
            rec['co_%s_attr' % tag] = c_at.numpy()
        u_ei, u_at = to_undirected(e.clone(), a.clone(), N)
        rec['und_ei'], rec['und_attr'] = u_ei.numpy(), u_at.numpy()
        rec['is_und'] = np.array([is_undirected(e.clone(), None, N), is_undirected(u_ei, None, N), is_undirected(u_ei, u_at, N)])
        np.savez_compressed(os.path.join(HERE, 'utils_rand_%s.npz' % name), **rec)
        print(name, 'E', E, 'unique', rec['co_row_ei'].shape[1])


if __name__ == '__main__':
    main()
