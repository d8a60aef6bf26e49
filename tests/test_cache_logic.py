"""CPU test of the cache rules in gammagl_b200/csr.py (no GPU, no kernels): identity keys with version counters, the exact content
match for rebuilt edge lists, entries of tensors modified in place, inference-mode tensors, LRU eviction."""
import torch

from gammagl_b200 import csr


def make_cache(capacity=4):
    c = csr._Lru(capacity, by_content=True)
    c.match_cpu = True
    return c


def test_identity_hit_and_content_hit():
    c = make_cache()
    built = []
    ei = torch.randint(0, 50, (2, 300))
    v1 = c.get(ei, 'g', lambda: built.append(1) or object())
    assert c.get(ei, 'g', lambda: built.append(1) or object()) is v1 and built == [1]          # same tensor: identity hit
    assert c.get(ei.clone(), 'g', lambda: built.append(1) or object()) is v1 and built == [1]  # new storage, equal contents
    assert c.content_hits == 1
    assert c.get(ei.clone(), 'other extra', lambda: built.append(1) or object()) is not v1      # extra (num_nodes ...) is part of the key
    other = ei.clone(); other[0, 0] = (other[0, 0] + 1) % 50
    assert c.get(other, 'g', lambda: built.append(1) or object()) is not v1                      # different contents: rebuilt
    assert len(built) == 3


def test_entry_of_a_tensor_modified_in_place_is_never_matched():
    c = make_cache()
    ei = torch.randint(0, 50, (2, 300))
    v1 = c.get(ei, 'g', object)
    ei[1, 0] = (ei[1, 0] + 7) % 50                     # version counter moves on: the cached value describes the OLD contents
    v2 = c.get(ei.clone(), 'g', object)                # equal to the NEW contents of `ei`, whose entry is stale
    assert v2 is not v1
    assert c.get(ei, 'g', object) is not v1            # identity lookup with the new version misses as well


def test_inference_tensors_are_not_cached_and_lru_evicts():
    c = make_cache(capacity=2)
    with torch.inference_mode():
        t = torch.arange(10)
    a, b = c.get(t, 'g', object), c.get(t, 'g', object)
    assert a is not b and len(c.d) == 0
    xs = [torch.full((4,), i) for i in range(3)]
    vals = [c.get(x, 'g', object) for x in xs]
    assert len(c.d) == 2 and c.get(xs[2], 'g', object) is vals[2]
    assert c.get(xs[0], 'g', object) is not vals[0]    # evicted, rebuilt


def test_derived_of_cpu_and_clear():
    csr.clear_caches()
    ei = torch.randint(0, 9, (2, 20))
    made = []
    a = csr.derived_of(ei, 'tag', lambda: made.append(1) or ('n', ei + 1))
    b = csr.derived_of(ei, 'tag', lambda: made.append(1) or ('n', ei + 1))
    assert a is b and made == [1] and len(csr._derived.d) == 1
    csr.clear_caches()
    assert len(csr._derived.d) == 0
