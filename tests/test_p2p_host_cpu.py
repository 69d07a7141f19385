"""Host-side bookkeeping of the peer mappings (api.p2p_open / p2p_release): an exported allocation
may be opened only once per process, several buffers can live in one allocation, and the mapping must
be closed exactly once, after its last user.  The CUDA calls are replaced by a recording stand-in."""
import ctypes as C

import pytest

from lattice_boltzmann_method_gpu_b200 import api


class FakeLib:
    def __init__(self, fail_on=None):
        self.opened, self.closed, self.next_ptr, self.fail_on = [], [], 0x7000_0000_0000, fail_on
        self.live = {}  # pointer -> handle of the mappings that are open right now

    def lbm_p2p_open(self, buf, out_ptr):
        handle = bytes(buf)
        if handle == self.fail_on:
            return -5
        assert handle not in self.live.values(), "an allocation must not be mapped twice at the same time"
        self.opened.append(handle)
        self.live[self.next_ptr] = handle
        out_ptr._obj.value = self.next_ptr
        self.next_ptr += 1 << 21
        return 0

    def lbm_p2p_close(self, ptr):
        assert ptr.value in self.live, "closing a mapping that is not open"
        del self.live[ptr.value]
        self.closed.append(ptr.value)
        return 0

    def lbm_last_error(self, _):
        return b"cudaIpcOpenMemHandle failed (stand-in)"


@pytest.fixture
def fake(monkeypatch):
    lib = FakeLib()
    monkeypatch.setattr(api, "load_library", lambda: lib)
    monkeypatch.setattr(api, "_opened_ipc", {})
    return lib


def test_same_allocation_is_opened_once_and_closed_after_the_last_release(fake):
    h1, h2 = b"A" * 64, b"B" * 64
    p1 = api.p2p_open(h1)
    assert api.p2p_open(h1) == p1          # second buffer of the same allocation: cached mapping
    p2 = api.p2p_open(h2)
    assert p2 != p1 and fake.opened == [h1, h2]
    api.p2p_release(h1)
    assert fake.closed == []               # still referenced once
    api.p2p_release(h1)
    assert fake.closed == [p1]
    api.p2p_release(h1)                    # unknown by now: ignored, no double close
    assert fake.closed == [p1]
    p1b = api.p2p_open(h1)                 # a later case may map the allocation again
    assert fake.opened == [h1, h2, h1] and p1b != p1
    api.p2p_release(h2), api.p2p_release(h1)
    assert fake.closed == [p1, p2, p1b] and api._opened_ipc == {}


def test_failed_open_raises_and_leaves_no_entry(monkeypatch):
    lib = FakeLib(fail_on=b"X" * 64)
    monkeypatch.setattr(api, "load_library", lambda: lib)
    monkeypatch.setattr(api, "_opened_ipc", {})
    with pytest.raises(api.LbmError) as e:
        api.p2p_open(b"X" * 64)
    assert e.value.status == -5 and api._opened_ipc == {}
