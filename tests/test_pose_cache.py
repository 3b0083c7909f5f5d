"""Host logic of the renderer's pose-class memo (lsi/geometry/ldi.py::_PoseClassCache): which hint (`variant` 5 = every image
rectified, 6 = none) a later forward_splat call may pass, decided only from flags the DEVICE computed for the same camera tensors.
No GPU needed: the state machine only looks at tensor identity and version counters."""
import torch

from lsi.geometry import ldi


class _DoneEvent(object):
    def query(self):
        return True


def _cams():
    return tuple(torch.zeros(2, 3, 3) for _ in range(3)) + (torch.zeros(2, 3, 1),)


def test_first_sight_gives_no_hint_and_flags_resolve_later():
    c = ldi._PoseClassCache()
    cams = _cams()
    e, cls = c.lookup(cams)
    assert cls is None and e['state'] is None
    e['pending'] = (torch.tensor([1, 1], dtype=torch.int32), _DoneEvent())      # what record() leaves once the copy has completed
    assert c.lookup(cams)[1] == 'all'
    e2, _ = c.lookup(_cams())
    e2['pending'] = (torch.tensor([0, 0], dtype=torch.int32), _DoneEvent())
    c2 = ldi._PoseClassCache()
    cams2 = _cams()
    e3, _ = c2.lookup(cams2)
    e3['pending'] = (torch.tensor([0, 1], dtype=torch.int32), _DoneEvent())
    assert c2.lookup(cams2)[1] == 'mixed'
    cams3 = _cams()
    e4, _ = c2.lookup(cams3)
    e4['pending'] = (torch.tensor([0, 0], dtype=torch.int32), _DoneEvent())
    assert c2.lookup(cams3)[1] == 'none'


def test_in_place_update_drops_the_hint_and_marks_the_set_volatile():
    c = ldi._PoseClassCache()
    cams = _cams()
    e, _ = c.lookup(cams)
    e['pending'] = (torch.tensor([1, 1], dtype=torch.int32), _DoneEvent())
    assert c.lookup(cams)[1] == 'all'
    cams[2].add_(1.0)                                   # version counter moves
    e, cls = c.lookup(cams)
    assert cls is None and e['volatile'] and e['state'] is None
    # a volatile set is never read back again (record() returns early), so it never gets a hint
    ldi._PoseClassCache.record(e, None, 2)
    assert e['pending'] is None and c.lookup(cams)[1] is None


def test_entries_are_tied_to_the_tensor_objects():
    c = ldi._PoseClassCache()
    cams = _cams()
    e, _ = c.lookup(cams)
    e['pending'] = (torch.tensor([1, 1], dtype=torch.int32), _DoneEvent())
    assert c.lookup(cams)[1] == 'all'
    key = tuple(id(x) for x in cams)
    other = _cams()
    c.entries[tuple(id(x) for x in other)] = c.entries.pop(key)      # as if the ids had been recycled by other tensors
    assert c.lookup(other)[1] is None


def test_content_key_entries():
    c = ldi._PoseClassCache()
    e, cls = c.lookup_key(('key', 123))
    assert cls is None
    e['pending'] = (torch.tensor([1], dtype=torch.int32), _DoneEvent())
    assert c.lookup_key(('key', 123))[1] == 'all'
    assert c.lookup_key(('key', 124))[1] is None
