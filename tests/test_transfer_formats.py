"""The transfer formats of include/kt_b200.h (compact: 32-bit label codes; packed: 16-bit label-pair indices) checked on
the CPU: the packers are exact or refuse, and the numpy mirror of the device expansion gives back the wide columns."""
import numpy as np
import pytest

from kube_throttler_b200 import abi, synth

FIELDS = ("labels", "req", "present", "flags", "ns_id")


@pytest.mark.parametrize("kw", [dict(config="C3", m=300, n=6000, p=800), dict(config="C2", m=200, n=5000, p=700, L=12),
                                dict(config="C2", m=40, n=70, p=33, R=1, L=3), dict(config="C4", m=500, n=3000, p=400), dict(config="C1")])
def test_packed_round_trip(kw):
    kw = dict(kw)
    snap = synth.generate(kw.pop("config"), **kw)
    for pods in (snap.running, snap.pending):
        pk = abi.packed_pods(pods)
        back = pk.unpack()
        for f in FIELDS:
            assert np.array_equal(getattr(back, f), getattr(pods, f)), f
        assert pk.labels16.dtype == np.uint16 and pk.meta.dtype == np.uint32 and pk.req32.dtype == np.int32
        assert pk.ns_bits + 3 + pods.req.shape[0] <= 32
        coded = abi.packed_pods(pods, code_requests=True)  # dictionary-coded request columns: same rows, fewer bytes
        back = coded.unpack()
        for f in FIELDS:
            assert np.array_equal(getattr(back, f), getattr(pods, f)), f
        assert coded.req32 is None and coded.req_codes.dtype == np.uint8 and coded.req_codes.shape[0] % 4 == 0
        assert coded.nbytes <= pk.nbytes + 8 * int(coded.req_dict.shape[0]) + 64


def test_coded_requests_pick_the_code_width_per_column():
    pods = synth.generate("C2", m=20, n=3000, p=5).running
    pods.req[1] = np.arange(3000, dtype=np.int64) * 7 - 11   # 3000 distinct values, negatives included: 2-byte codes
    pk = abi.packed_pods(pods, code_requests=True)
    assert list(pk.req_code_bytes) == [1, 2, 1, 1]
    assert np.array_equal(pk.unpack().req, pods.req)
    pods.req[2, :] = 0                                        # a constant column: one dictionary entry
    assert np.array_equal(abi.packed_pods(pods, code_requests=True).unpack().req, pods.req)
    big = synth.generate("C2", m=10, n=70000, p=3).running
    big.req[0] = np.arange(70000, dtype=np.int64)
    with pytest.raises(ValueError):
        abi.packed_pods(big, code_requests=True)


def test_packed_refuses_what_it_cannot_carry():
    pods = synth.generate("C2", m=40, n=70, p=33).running
    pods.req[0, 3] = (1 << 40) + 1  # odd and huge: no power-of-two unit makes the column fit int32
    with pytest.raises(ValueError):
        abi.packed_pods(pods)
    pods = synth.generate("C2", m=40, n=70, p=33).running
    pods.ns_id[0] = 1 << 26          # R=4: 25 namespace bits
    with pytest.raises(ValueError):
        abi.packed_pods(pods)
    big = synth.generate("C2", m=10, n=70000, p=3).running
    big.labels[0, :] = (np.int64(7) << 32) | np.arange(70000, dtype=np.int64)  # 70000 distinct pairs
    with pytest.raises(ValueError):
        abi.packed_pods(big)


def test_packed_is_a_third_of_the_wide_rows():
    pods = synth.generate("C2", m=50, n=20000, p=10).running
    wide = sum(getattr(pods, f).nbytes for f in FIELDS)
    assert abi.packed_pods(pods).nbytes < 0.36 * wide < abi.compact_pods(pods).nbytes
