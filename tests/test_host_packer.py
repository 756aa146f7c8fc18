"""Host-side packer of the PRODUCT (kube_throttler_b200/csrc/kt_host.cc via kth_eval), CPU only.

The same vectors that pin the oracle (reference unit tests) are run against the product's own
Quantity parser, PodRequestResourceList, RFC3339/override messages and selector validation, and the two
independent implementations (product host layer vs oracle) are compared on a wider input set.
No device is touched: kth_eval is host-only by construction.
"""
from fractions import Fraction

import pytest

from test_oracle_kat import rl_values


@pytest.fixture(scope="module")
def host(kt):
    from kube_throttler_b200 import host as h

    return h


QUANTITIES = [
    ("0", 0), ("1", 1), ("-2", -2), ("+3", 3), ("500m", Fraction(1, 2)), ("1.1", Fraction(11, 10)), ("0.200", Fraction(1, 5)),
    ("512Mi", 536870912), ("1Gi", 2**30), ("1Ki", 1024), ("1.5Gi", 3 * 2**29), ("100n", Fraction(1, 10**7)), ("5u", Fraction(5, 10**6)),
    ("3k", 3000), ("2M", 2 * 10**6), ("1G", 10**9), ("1T", 10**12), ("1P", 10**15), ("1E", 10**18), ("1e3", 1000), ("1E-3", Fraction(1, 1000)),
    ("12e6", 12 * 10**6), (".5", Fraction(1, 2)), ("5.", 5), ("0.0000000001", Fraction(1, 10**9)), ("1.0000000001", Fraction(10**9 + 1, 10**9)),
    ("-0.0000000001", Fraction(-1, 10**9)), ("16Ei", 2**63 - 1), ("0.1Ki", Fraction(1024, 10)), ("50m", Fraction(1, 20)), ("1000m", 1),
]


@pytest.mark.parametrize("s,want", QUANTITIES)
def test_parse_quantity_values(host, oracle, s, want):
    got = host.eval_host("ParseQuantity", value=s)
    assert Fraction(got["decimal"]) == want
    ref = oracle.call("ParseQuantity", value=s)  # two independent parsers agree, value and format class
    assert Fraction(ref["decimal"]) == want and ref["format"] == got["format"]


@pytest.mark.parametrize("s", ["", "abc", "1x", "1Kii", "--1", "1e", "1e1.5", "1 ", " 1", "1.1.M", "0.1mi", "1i"])
def test_parse_quantity_errors(host, s):
    with pytest.raises(RuntimeError):
        host.eval_host("ParseQuantity", value=s)


@pytest.mark.parametrize("s,canon", [("500m", "500m"), ("0.5", "500m"), ("1", "1"), ("1000m", "1"), ("900m", "900m"), ("1.1", "1100m"),
                                     ("512Mi", "512Mi"), ("1536Mi", "1536Mi"), ("1Gi", "1Gi"), ("2000", "2k"), ("100n", "100n"), ("0", "0")])
def test_canonical_quantity_strings(host, s, canon):
    """resource.Quantity.String spellings the integration suite asserts on status fields
    (util_throttle_test.go:169-177: "500m", "200m", "900m", "1"; README.md:287-309: 512Mi)."""
    assert host.eval_host("CanonicalQuantity", value=s)["canonical"] == canon


def test_pod_request_resource_list(host, oracle):
    """resourcelist_test.go:47-117"""
    pod = {"kind": "Pod", "metadata": {"name": "p"}, "spec": {"containers": [
        {"resources": {"requests": {"n1": "1"}}}, {"resources": {"requests": {"n1": "1"}}}]}}
    assert rl_values(host.eval_host("PodRequestResourceList", pod=pod)) == {"n1": 2}
    pod["spec"]["initContainers"] = [{"resources": {"requests": {"n1": "1"}}}, {"resources": {"requests": {"n2": "2"}}}]
    assert rl_values(host.eval_host("PodRequestResourceList", pod=pod)) == {"n1": 2, "n2": 2}
    pod["spec"]["overhead"] = {"n1": "500m", "n3": "1"}
    assert rl_values(host.eval_host("PodRequestResourceList", pod=pod)) == {"n1": Fraction(5, 2), "n2": 2, "n3": 1}
    # init container larger than the container sum wins; zero-valued init-only names are still inserted (SetMax, :76-84)
    pod = {"spec": {"initContainers": [{"resources": {"requests": {"cpu": "3", "x": "0"}}}],
                    "containers": [{"resources": {"requests": {"cpu": "1"}}}, {"resources": {"requests": {"cpu": "500m", "memory": "1Gi"}}}]}}
    want = {"cpu": 3, "x": 0, "memory": 2**30}
    assert rl_values(host.eval_host("PodRequestResourceList", pod=pod)) == want
    assert rl_values(oracle.call("PodRequestResourceList", pod=pod)) == want
    amt = host.eval_host("ResourceAmountOfPod", pod=pod)  # resource_amount.go:71-76
    assert amt["resourceCounts"] == {"pod": 1} and rl_values(amt["resourceRequests"]) == want


@pytest.mark.parametrize("s", ["2026-01-01T00:00:00Z", "2019-02-01T00:00:00+09:00", "2021-08-04T10:00:00.5Z", "error", "2021-13-04T10:00:00Z",
                               "2021-02-30T10:00:00Z", "2021-02-03", "2021-02-03T10:00:00", "2021-02-03T10:00:00Zjunk", "2021-02-03T25:00:00Z",
                               "2021-02-03T10:00:00+0900", "", "20210203T100000Z"])
def test_rfc3339_matches_oracle(host, oracle, s):
    """time.Parse(time.RFC3339, s): value and Go's exact error text (throttle_types_test.go:147 pins one of them)."""
    got, want = host.eval_host("ParseRFC3339", value=s), oracle.call("ParseRFC3339", value=s)
    assert got.get("error") == want.get("error")
    if "error" not in want:
        assert (got["unix"], got["nsec"]) == (want["unix"], want["nsec"])


def test_override_messages(host):
    """throttle_types_test.go:110-152: unparsable overrides are skipped and reported by index"""
    thr = {"kind": "Throttle", "metadata": {"name": "t", "namespace": "default"},
           "spec": {"throttlerName": "dummy", "threshold": {"resourceRequests": {"cpu": "1"}}, "temporaryThresholdOverrides": [
               {"begin": "2006-01-02T15:03:05Z", "end": "2006-01-02T15:05:05Z", "threshold": {"resourceRequests": {"cpu": "2"}}},
               {"begin": "error", "end": "error"}, {"begin": "", "end": "nope"}]}}
    assert host.eval_host("OverrideMessages", throttle=thr) == [
        'index 1: Failed to parse Begin: parsing time "error" as "2006-01-02T15:04:05Z07:00": cannot parse "error" as "2006"',
        'index 2: Failed to parse End: parsing time "nope" as "2006-01-02T15:04:05Z07:00": cannot parse "nope" as "2006"']


@pytest.mark.parametrize("now", ["2021-08-04T09:00:00Z", "2021-08-04T10:00:00Z", "2021-08-04T12:00:00+09:00", "2021-08-05T10:00:00Z", "2021-08-06T00:00:00Z"])
def test_next_override_happens_in(host, oracle, now):
    """throttle_types.go:37-63: nearest begin/end after now; entries whose Begin does not parse are skipped whole."""
    thr = {"kind": "Throttle", "metadata": {"name": "t", "namespace": "default"},
           "spec": {"throttlerName": "x", "threshold": {}, "temporaryThresholdOverrides": [
               {"begin": "2021-08-04T10:00:00Z", "end": "2021-08-05T10:00:00Z"}, {"begin": "garbage", "end": "2021-08-04T09:30:00Z"},
               {"begin": "", "end": "2021-08-05T12:00:00Z"}, {"begin": "2021-08-04T11:00:00Z", "end": "nonsense"}]}}
    got, want = host.eval_host("NextOverrideHappensIn", throttle=thr, now=now), oracle.call("NextOverrideHappensIn", throttle=thr, now=now)
    assert got["have"] == want["have"] and (not want["have"] or got["nanos"] == want["nanos"])


@pytest.mark.parametrize("sel,valid", [
    ({}, True), ({"matchLabels": {"a": "b"}}, True), ({"matchExpressions": [{"key": "a", "operator": "In", "values": ["x"]}]}, True),
    ({"matchExpressions": [{"key": "a", "operator": "Exists"}]}, True), ({"matchExpressions": [{"key": "a", "operator": "In", "values": []}]}, False),
    ({"matchExpressions": [{"key": "a", "operator": "Exists", "values": ["x"]}]}, False), ({"matchExpressions": [{"key": "a", "operator": "Bogus"}]}, False),
    ({"matchLabels": {"bad key!": "b"}}, False), ({"matchLabels": {"a": "bad value!"}}, False), ({"matchLabels": {"example.com/role": "db"}}, True),
    ({"matchLabels": {"a": "x" * 64}}, False), ({"matchLabels": {"/name": "x"}}, False),
])
def test_selector_validation(host, sel, valid):
    """metav1.LabelSelectorAsSelector / labels.NewRequirement validation (PARITY UNPINNED by reference tests; apimachinery v0.26.4 rules)"""
    assert host.eval_host("ValidateSelector", selector=sel)["valid"] is valid


def test_new_plugin_needs_a_gpu_or_fails(host, kt):
    """NewPlugin without a device is an error, never a CPU fallback; bad args are rejected like DecodePluginArgs."""
    import ctypes as C
    L = host._bind()
    h = C.c_void_p()
    assert L.kth_new_plugin(C.byref(h), b'{"targetSchedulerName":"s"}', 0) == kt.abi.ERR_INVALID and not h
    assert L.kth_new_plugin(C.byref(h), b'{"name":"n"}', 0) == kt.abi.ERR_INVALID and not h
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(kt.KtError) as e:
            host.Plugin()
        assert e.value.code == kt.abi.ERR_CUDA


# ---- gauges (throttle_metrics.go, clusterthrottle_metrics.go, metrics_recorder.go) ----------------------------------------
# The reference has no test for its recorders (parity unpinned by reference vectors); these pin the recorder's rules as
# the code states them -- counts nil -> 0, cpu as MilliValue, other resources as Value (both round up), label sets, family
# names -- and Go's float spelling ('g', shortest) that the Prometheus text format uses.

@pytest.mark.parametrize("value,text", [
    ("0", "0"), ("1", "1"), ("-1", "-1"), ("2", "2"), ("200", "200"), ("999999", "999999"), ("1000000", "1e+06"), ("1234567", "1.234567e+06"),
    ("536870912", "5.36870912e+08"), ("1073741824", "1.073741824e+09"), ("1.5", "1.5"), ("0.0001", "0.0001"), ("0.00001", "1e-05"),
    ("-2.5", "-2.5"), ("1e21", "1e+21"), ("9007199254740993", "9.007199254740992e+15"),
])
def test_gauge_value_spelling(host, value, text):
    assert host.eval_host("FormatFloat", value=value)["text"] == text


@pytest.mark.parametrize("q,scale,want", [
    ("500m", -3, 500), ("1", -3, 1000), ("1.1", -3, 1100), ("100n", -3, 1), ("0.0001", -3, 1), ("512Mi", 0, 536870912), ("1500m", 0, 2),
    ("0.5", 0, 1), ("-1500m", 0, -2), ("0", 0, 0), ("3", 0, 3), ("1Ei", 0, 2**60),
])
def test_scaled_value_rounds_up(host, q, scale, want):
    """Quantity.MilliValue / Value = ScaledValue(-3 / 0): rounded away from zero (apimachinery v0.26.4 amount.go)."""
    assert host.eval_host("ScaledValue", value=q, scale=scale)["value"] == want


def _series(text):
    out = {}
    for line in text.splitlines():
        if line and not line.startswith("#"):
            k, v = line.rsplit(" ", 1)
            out[k] = v
    return out


def test_throttle_metrics_of_one_object(host):
    thr = {"kind": "Throttle", "metadata": {"name": "t1", "namespace": "default", "uid": "u-1"},
           "spec": {"throttlerName": "kube-throttler", "threshold": {"resourceCounts": {"pod": 5}, "resourceRequests": {"cpu": "200m", "memory": "1Gi"}}},
           "status": {"calculatedThreshold": {"threshold": {"resourceRequests": {"cpu": "700m"}}, "calculatedAt": "2026-01-01T00:00:00Z"},
                      "throttled": {"resourceCounts": {"pod": False}, "resourceRequests": {"cpu": True}},
                      "used": {"resourceCounts": {"pod": 3}, "resourceRequests": {"cpu": "1500m", "memory": "512Mi", "example.com/dongle": "2500m"}}}}
    text = host.eval_host("ThrottleMetrics", throttle=thr)["text"]
    s = _series(text)
    lab = lambda r: '{name="t1",namespace="default",resource="%s",uid="u-1"}' % r  # noqa: E731  (label pairs sorted by name)
    assert s["throttle_spec_threshold_resourceCounts" + lab("pod")] == "5"
    assert s["throttle_spec_threshold_resourceRequests" + lab("cpu")] == "200"            # MilliValue
    assert s["throttle_spec_threshold_resourceRequests" + lab("memory")] == "1.073741824e+09"
    assert s["throttle_status_calculated_threshold_resourceCounts" + lab("pod")] == "0"    # nil counts record 0
    assert s["throttle_status_calculated_threshold_resourceRequests" + lab("cpu")] == "700"
    assert "throttle_status_calculated_threshold_resourceRequests" + lab("memory") not in s
    assert s["throttle_status_throttled_resourceCounts" + lab("pod")] == "0"
    assert s["throttle_status_throttled_resourceRequests" + lab("cpu")] == "1"
    assert s["throttle_status_used_resourceCounts" + lab("pod")] == "3"
    assert s["throttle_status_used_resourceRequests" + lab("cpu")] == "1500"
    assert s["throttle_status_used_resourceRequests" + lab("memory")] == "5.36870912e+08"
    assert s["throttle_status_used_resourceRequests" + lab("example.com/dongle")] == "3"  # Value() rounds 2.5 up
    assert "# TYPE throttle_status_used_resourceRequests gauge" in text
    assert "# HELP throttle_status_used_resourceCounts used resource counts of the throttle" in text
    fams = [l.split()[2] for l in text.splitlines() if l.startswith("# TYPE")]
    assert fams == sorted(fams) and len(fams) == 8


def test_clusterthrottle_metrics_have_no_namespace_label(host):
    thr = {"kind": "ClusterThrottle", "metadata": {"name": 'c"1', "uid": "u-2"},
           "spec": {"throttlerName": "kube-throttler", "threshold": {"resourceRequests": {"nvidia.com/gpu": "8"}}}}
    s = _series(host.eval_host("ThrottleMetrics", throttle=thr)["text"])
    assert s['clusterthrottle_spec_threshold_resourceRequests{name="c\\"1",resource="nvidia.com/gpu",uid="u-2"}'] == "8"
    assert s['clusterthrottle_spec_threshold_resourceCounts{name="c\\"1",resource="pod",uid="u-2"}'] == "0"
    assert not any(k.startswith("throttle_") for k in s)
    assert not any("resourceRequests" in k and "status" in k for k in s)  # nil maps record nothing


@pytest.mark.parametrize("s,want", [
    (".", 0), ("-", 0), ("+", 0), ("-.", 0), (".m", 0), (".Ki", 0), ("m", 0), ("Ki", 0), ("e3", 0), ("00", 0), ("0.0", 0), ("-0", 0),
    ("000.500", Fraction(1, 2)), ("1.", 1), ("1.G", 10**9), ("5.m", Fraction(1, 200)), ("1E", 10**18), ("12E", 12 * 10**18), ("+1e3", 1000),
])
def test_degenerate_quantity_spellings(host, oracle, s, want):
    """parseQuantityString (apimachinery v0.26.4 quantity.go) replaces an empty numerator by "0" and allows an empty
    denominator ("we currently allow 1.G"), so these all parse.  Product and oracle, written independently, agree."""
    assert Fraction(host.eval_host("ParseQuantity", value=s)["decimal"]) == want
    assert Fraction(oracle.call("ParseQuantity", value=s)["decimal"]) == want


@pytest.mark.parametrize("s", ["1Ee", "1ee3", "1e3e", "1.5.5", "..", "1,5", "1 m", "0x10"])
def test_malformed_quantities_rejected_by_both(host, oracle, s):
    with pytest.raises(RuntimeError):
        host.eval_host("ParseQuantity", value=s)
    with pytest.raises(RuntimeError):
        oracle.call("ParseQuantity", value=s)


# ---- randomised differential checks: the product's packer and the oracle were written independently ----------------------

def _rand_quantity(rng):
    kind = rng.random()
    if kind < 0.35:
        s = str(rng.randrange(0, 5000)) + rng.choice(["", "m", "m", "k", "Ki", "Mi", "Gi", "M", "G", "u", "n"])
    elif kind < 0.7:
        s = f"{rng.randrange(0, 400)}.{rng.randrange(0, 10 ** rng.randrange(1, 5)):0{rng.randrange(1, 5)}d}" + rng.choice(["", "", "m", "Ki", "Mi", "Gi", "k"])
    elif kind < 0.85:
        s = f"{rng.randrange(1, 999)}e{rng.randrange(-6, 7)}"
    else:
        s = rng.choice(["0", "00", "0.0", ".5", "5.", "1e0", "+3", "-2", "-1500m", "100n", "1n", "0.0000000004", "1.0000000005", "15Ei", "8191Pi"])
    return s


@pytest.mark.parametrize("seed", range(8))
def test_random_quantities_agree_with_the_oracle(host, oracle, seed):
    import random

    rng = random.Random(1000 + seed)
    for _ in range(250):
        s = _rand_quantity(rng)
        a, b = host.eval_host("ParseQuantity", value=s), oracle.call("ParseQuantity", value=s)
        assert Fraction(a["decimal"]) == Fraction(b["decimal"]), s
        assert a["format"] == b["format"], s
        canon = host.eval_host("CanonicalQuantity", value=s)["canonical"]  # re-parsing the canonical spelling gives the value back
        assert Fraction(host.eval_host("ParseQuantity", value=canon)["decimal"]) == Fraction(a["decimal"]), (s, canon)


@pytest.mark.parametrize("seed", range(6))
def test_random_pods_request_lists_agree_with_the_oracle(host, oracle, seed):
    """PodRequestResourceList over random container / initContainer / overhead shapes (resourcelist.go:27-46)."""
    import random

    rng = random.Random(2000 + seed)
    names = ["cpu", "memory", "nvidia.com/gpu", "ephemeral-storage", "example.com/x"]

    def reqs():
        return {n: _rand_quantity(rng).lstrip("-") or "0" for n in rng.sample(names, rng.randrange(0, 4))}

    for _ in range(60):
        spec = {"containers": [{"name": f"c{i}", "resources": {"requests": reqs()}} for i in range(rng.randrange(0, 4))]}
        if rng.random() < 0.5:
            spec["initContainers"] = [{"name": f"i{i}", "resources": {"requests": reqs()}} for i in range(rng.randrange(1, 3))]
        if rng.random() < 0.3:
            spec["overhead"] = reqs()
        pod = {"kind": "Pod", "metadata": {"name": "p", "namespace": "d"}, "spec": spec}
        got, want = rl_values(host.eval_host("PodRequestResourceList", pod=pod)), rl_values(oracle.call("PodRequestResourceList", pod=pod))
        assert got == want, pod


@pytest.mark.parametrize("seed", range(6))
def test_random_selectors_validate_like_the_oracle(host, oracle, seed):
    """LabelSelectorAsSelector accepts / rejects the same selectors in both implementations (operators, value counts, key and
    value syntax).  PARITY UNPINNED by reference tests; both follow apimachinery v0.26.4 validation rules independently."""
    import random

    rng = random.Random(3000 + seed)
    keys = ["app", "tier", "example.com/role", "a" * 63, "a" * 64, "-bad", "bad-", "a/b/c", "/x", "Ex_Ample.com/k", "exa_mple.com/k", "k.", "", "x y"]
    vals = ["db", "", "a" * 63, "a" * 64, "bad value", "-x", "x-", "x_y.z", "9"]
    ops = ["In", "NotIn", "Exists", "DoesNotExist", "Bogus", "in", ""]
    disagreements = []
    for _ in range(150):
        sel = {}
        if rng.random() < 0.6:
            sel["matchLabels"] = {rng.choice(keys): rng.choice(vals) for _ in range(rng.randrange(0, 3))}
        if rng.random() < 0.7:
            sel["matchExpressions"] = [{"key": rng.choice(keys), "operator": rng.choice(ops), "values": [rng.choice(vals) for _ in range(rng.randrange(0, 3))]}
                                       for _ in range(rng.randrange(0, 3))]
        got = host.eval_host("ValidateSelector", selector=sel)
        ref = oracle.call("ThrottleSelector.MatchesToPod", selector={"selectorTerms": [{"podSelector": sel}]},
                          pod={"metadata": {"name": "p", "namespace": "d", "labels": {"app": "db"}}, "spec": {}})
        if got["valid"] != ("error" not in ref) or got.get("error") != ref.get("error"):  # the same verdict AND the same message
            disagreements.append((sel, got, ref))
    assert not disagreements, disagreements[:3]


@pytest.mark.parametrize("sel,message", [
    ({"matchExpressions": [{"key": "a", "operator": "Exists", "values": ["x", "y"]}]},
     'values: Invalid value: []string{"x", "y"}: values set must be empty for exists and does not exist'),
    ({"matchExpressions": [{"key": "a", "operator": "In", "values": []}]},
     "values: Invalid value: []string(nil): for 'in', 'notin' operators, values set can't be empty"),
    ({"matchExpressions": [{"key": "a", "operator": "Bogus"}]}, '"Bogus" is not a valid label selector operator'),
    ({"matchLabels": {"bad key!": "b"}},
     "key: Invalid value: \"bad key!\": name part must consist of alphanumeric characters, '-', '_' or '.', and must start and end with an alphanumeric "
     "character (e.g. 'MyName',  or 'my.name',  or '123-abc', regex used for validation is '([A-Za-z0-9][-A-Za-z0-9_.]*)?[A-Za-z0-9]')"),
    ({"matchLabels": {"a": "x" * 64}}, 'values[0][a]: Invalid value: "' + "x" * 64 + '": must be no more than 63 characters'),
    ({"matchLabels": {"/name": "x"}}, 'key: Invalid value: "/name": prefix part must be non-empty'),
])
def test_selector_error_messages(host, oracle, sel, message):
    """The text of labels.NewRequirement's error as apimachinery v0.26.4 words it (field.Invalid with %q / %#v of the value,
    validation.RegexError's "(e.g. ...)" tail): PreFilter's Error status carries it verbatim (plugin.go:154-156).  PARITY
    UNPINNED by reference tests; product and oracle restate it independently and must agree."""
    assert host.eval_host("ValidateSelector", selector=sel)["error"] == message
    ref = oracle.call("ThrottleSelector.MatchesToPod", selector={"selectorTerms": [{"podSelector": sel}]},
                      pod={"metadata": {"name": "p", "namespace": "d", "labels": {}}, "spec": {}})
    assert ref["error"] == message


def test_random_timestamps_parse_like_the_oracle(host, oracle):
    """time.Parse(time.RFC3339, s) over well-formed and malformed spellings (month / day / hour ranges, leap days, fractional
    seconds with '.' and ',', zone offsets with and without colon, lower-case separators): value and Go's error text agree."""
    import random

    rng = random.Random(7)
    for _ in range(600):
        y = rng.choice([1, 1969, 1970, 1999, 2000, 2021, 2024, 2026, 2100, 9999])
        mo, d = rng.choice([0, 1, 2, 2, 2, 6, 12, 13]), rng.choice([0, 1, 28, 29, 30, 31, 32])
        h, mi, sec = rng.choice([0, 12, 23, 24, 25]), rng.choice([0, 30, 59, 60]), rng.choice([0, 30, 59, 60, 61])
        frac = rng.choice(["", "", ".5", ".123456789", ".1234567891", ".", ",5", ".000"])
        tz = rng.choice(["Z", "Z", "+09:00", "-07:00", "+00:00", "+24:00", "+09:60", "+0900", "z", "", " Z", "+9:00", "-00:00", "+23:59"])
        s = f"{y:04d}-{mo:02d}-{d:02d}{rng.choice('TTTt ')}{h:02d}:{mi:02d}:{sec:02d}{frac}{tz}"
        got, want = host.eval_host("ParseRFC3339", value=s), oracle.call("ParseRFC3339", value=s)
        assert got.get("error") == want.get("error"), s
        if "error" not in want:
            assert (got["unix"], got["nsec"]) == (want["unix"], want["nsec"]), s


def test_manifest_parser_refuses_pathological_json(host):
    """kth_* take manifests as JSON text: nesting is bounded (the parser recurses), malformed input is an error, not a crash."""
    L = host._bind()
    deep = b'{"fn":"PodRequestResourceList","pod":{"spec":{},"x":' + b"[" * 300 + b"]" * 300 + b"}}"
    assert b"nested too deeply" in L.kth_eval(deep)
    ok = b'{"fn":"PodRequestResourceList","pod":{"spec":{},"x":' + b"[" * 200 + b"]" * 200 + b"}}"
    assert L.kth_eval(ok) == b"{}"
    for bad in (b"", b"{", b'{"fn":', b'{"fn":"ParseQuantity","value":"1"} trailing', b'{"fn":"ParseQuantity","value":"\\u12"}', b"[" * 100000):
        out = L.kth_eval(bad)
        assert out.startswith(b'{"error"'), bad[:40]
