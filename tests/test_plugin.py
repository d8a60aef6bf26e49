"""CPU test of the host-side plug-in logic (gammagl_b200/plugin.py) against a stand-in `gammagl` package laid out
like the reference (mpops star-export copied into message_passing / utils modules, conv classes without
message_aggregate).  No compute is executed."""
import sys
import types

from gammagl_b200 import mpops, plugin


def _fake_gammagl():
    def orig(*a, **k):
        raise AssertionError('original operator must not run')
    pkg = types.ModuleType('gammagl')
    mp = types.ModuleType('gammagl.mpops')
    mpt = types.ModuleType('gammagl.mpops.torch')
    for m in (mp, mpt):
        for n in plugin._NAMES:
            setattr(m, n, orig)
        m.use_ext = False
    msg = types.ModuleType('gammagl.layers.conv.message_passing')
    for n in plugin._NAMES[:8]:
        setattr(msg, n, orig)                      # `from gammagl.mpops import *` copies (message_passing.py:2)
    msg.use_ext = False

    class MessagePassing:
        def propagate(self, *a, **k):
            return 'message_aggregate' in self.__class__.__dict__
    msg.MessagePassing = MessagePassing
    conv = types.ModuleType('gammagl.layers.conv')
    for n in plugin._FUSED_LAYERS + ['HGTConv']:
        setattr(conv, n, type(n, (MessagePassing,), {}))
    sm = types.ModuleType('gammagl.utils.softmax')
    sm.segment_softmax = orig
    sm.unsorted_segment_max = orig
    ut = types.ModuleType('gammagl.utils')
    ut.coalesce = lambda edge_index, *a, **k: ('original', edge_index)       # the function shadows the submodule name
    ut.sort_edge_index = lambda edge_index, *a, **k: ('original', edge_index)
    utc = types.ModuleType('gammagl.utils.coalesce')
    utc.coalesce = ut.coalesce
    mods_extra = {'gammagl.utils': ut, 'gammagl.utils.coalesce': utc}
    mods = {'gammagl': pkg, 'gammagl.mpops': mp, 'gammagl.mpops.torch': mpt,
            'gammagl.layers.conv.message_passing': msg, 'gammagl.layers.conv': conv, 'gammagl.utils.softmax': sm}
    mods.update(mods_extra)
    return mods, orig


def test_enable_rebinds_every_copy_and_disable_restores():
    mods, orig = _fake_gammagl()
    n = plugin.enable(modules=mods)
    assert n >= 2 * len(plugin._NAMES) + 8
    assert mods['gammagl.mpops'].gspmm is mpops.gspmm
    assert mods['gammagl.mpops.torch'].c_spmm_sum is mpops.c_spmm_sum
    assert mods['gammagl.mpops'].use_ext is True
    assert mods['gammagl.layers.conv.message_passing'].unsorted_segment_sum is mpops.unsorted_segment_sum
    assert mods['gammagl.utils.softmax'].segment_softmax is mpops.segment_softmax
    conv = mods['gammagl.layers.conv']
    for name in plugin._FUSED_LAYERS:
        assert getattr(conv, name)().propagate() is True      # reference switch (message_passing.py:144) now fuses
        assert getattr(conv, name).message_aggregate.__name__ == 'message_aggregate'   # Inspector files signatures by __name__
    assert conv.HGTConv().propagate() is False                # layers with their own message() are left alone
    # utilities: CPU tensors / numpy keep the original implementation, CUDA tensors would go to the GPU sort
    import torch
    cpu_ei = torch.tensor([[1, 0], [0, 1]])
    original_coalesce = mods['gammagl.utils.coalesce'].coalesce._gglb200_original
    assert mods['gammagl.utils'].coalesce(cpu_ei) == ('original', cpu_ei)
    assert mods['gammagl.utils.coalesce'].coalesce(cpu_ei)[0] == 'original'
    assert mods['gammagl.utils'].sort_edge_index([[0], [1]])[0] == 'original'
    assert plugin.enable(modules=mods) == 0                    # idempotent: nothing left to rebind
    plugin.disable()
    assert mods['gammagl.utils'].coalesce is original_coalesce
    assert mods['gammagl.mpops'].gspmm is orig
    assert mods['gammagl.mpops'].use_ext is False
    assert getattr(conv, 'GCNConv')().propagate() is False


def test_enable_without_gammagl_raises():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'gammagl' or k.startswith('gammagl.')}
    try:
        try:
            plugin.enable()
        except ImportError:
            pass
        else:
            raise AssertionError('enable() must raise when gammagl is not imported')
    finally:
        sys.modules.update(saved)
