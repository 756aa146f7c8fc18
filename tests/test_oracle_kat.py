"""Pins the ORACLE against the reference's own tests (known-answer vectors transcribed by hand;
each block names the reference test file:line it restates).  CPU only.

If one of these fails the oracle is wrong and no GPU parity claim means anything.
"""
from fractions import Fraction

import pytest


def q(s):
    """Numeric value of a quantity string as the oracle prints/accepts it (tests compare values, not spellings)."""
    suffixes = {"Ki": 2**10, "Mi": 2**20, "Gi": 2**30, "Ti": 2**40, "Pi": 2**50, "Ei": 2**60,
                "n": Fraction(1, 10**9), "u": Fraction(1, 10**6), "m": Fraction(1, 1000), "k": 10**3, "M": 10**6,
                "G": 10**9, "T": 10**12, "P": 10**15, "E": 10**18}
    s = str(s)
    for suf in sorted(suffixes, key=len, reverse=True):
        if s.endswith(suf):
            return Fraction(s[: -len(suf)]) * suffixes[suf]
    return Fraction(s)


def rl_values(d):
    return {k: q(v) for k, v in d.items()}


def mk_pod(name="test", namespace="test", labels=None, requests=None):
    """v1alpha1_suite_test.go:44-75 mkPod(...).WithLabels().WithRequests()"""
    ctr = {"name": "ctr", "image": "dummy"}
    if requests is not None:
        ctr["resources"] = {"requests": requests}
    md = {"name": name, "namespace": namespace}
    if labels is not None:
        md["labels"] = labels
    return {"kind": "Pod", "metadata": md, "spec": {"containers": [ctr]}}


def mk_ns(name, labels):
    return {"kind": "Namespace", "metadata": {"name": name, "labels": labels or {}}}


# ------------------------------------------------------------------------------------------------
# resource_amount_test.go:27-210  ResourceAmount.IsThrottled
# ------------------------------------------------------------------------------------------------
def is_throttled(oracle, threshold, used, on_equal):
    return oracle.call("ResourceAmount.IsThrottled", threshold=threshold, used=used, onEqual=on_equal)


def test_is_throttled_empty_threshold(oracle):
    for b in (False, True):
        r = is_throttled(oracle, {}, {"resourceCounts": {"pod": 3}}, b)
        assert r == {"resourceCounts": {"pod": False}}
        r = is_throttled(oracle, {}, {"resourceRequests": {"r1": "1000"}}, b)
        assert r == {"resourceCounts": {"pod": False}}  # ResourceRequests stays nil


THRESHOLD = {"resourceCounts": {"pod": 3}, "resourceRequests": {"r1": "10", "r2": "20"}}
RR_FALSE = {"r1": False, "r2": False}


def test_is_throttled_counts(oracle):
    for b in (False, True):
        assert is_throttled(oracle, THRESHOLD, {"resourceCounts": {"pod": 2}}, b) == {
            "resourceCounts": {"pod": False}, "resourceRequests": RR_FALSE}
    assert is_throttled(oracle, THRESHOLD, {"resourceCounts": {"pod": 3}}, False)["resourceCounts"] == {"pod": False}
    assert is_throttled(oracle, THRESHOLD, {"resourceCounts": {"pod": 3}}, True)["resourceCounts"] == {"pod": True}
    for b in (False, True):
        assert is_throttled(oracle, THRESHOLD, {"resourceCounts": {"pod": 4}}, b) == {
            "resourceCounts": {"pod": True}, "resourceRequests": RR_FALSE}


@pytest.mark.parametrize("used,on_equal,want", [
    ({"r1": "1", "r2": "2"}, False, {"r1": False, "r2": False}),
    ({"r1": "1", "r2": "2"}, True, {"r1": False, "r2": False}),
    ({"r1": "10", "r2": "20"}, False, {"r1": False, "r2": False}),
    ({"r1": "10", "r2": "20"}, True, {"r1": True, "r2": True}),
    ({"r1": "11", "r2": "22"}, False, {"r1": True, "r2": True}),
    ({"r1": "11", "r2": "22"}, True, {"r1": True, "r2": True}),
    ({"r1": "1", "r2": "20"}, False, {"r1": False, "r2": False}),
    ({"r1": "1", "r2": "20"}, True, {"r1": False, "r2": True}),
    ({"r3": "3000"}, False, {"r1": False, "r2": False}),  # resources only in `used` are ignored
    ({"r3": "3000"}, True, {"r1": False, "r2": False}),
])
def test_is_throttled_requests(oracle, used, on_equal, want):
    r = is_throttled(oracle, THRESHOLD, {"resourceRequests": used}, on_equal)
    assert r == {"resourceCounts": {"pod": False}, "resourceRequests": want}


# resource_amount_test.go:212-250  IsResourceAmountThrottled.IsThrottledFor
def test_is_throttled_for(oracle):
    f = lambda thr, pod: oracle.call("IsThrottledFor", throttled=thr, pod=pod)
    assert f({"resourceCounts": {"pod": True}}, mk_pod()) is True
    t = {"resourceRequests": {"r1": False, "r2": True}}
    assert f(t, mk_pod(requests={"r2": "0"})) is False   # zero request: present but skipped (Q5)
    assert f(t, mk_pod(requests={"r2": "1"})) is True
    assert f(t, mk_pod(requests={"r1": "1000"})) is False
    assert f(t, mk_pod(requests={"r3": "1000"})) is False


# ------------------------------------------------------------------------------------------------
# throttle_types_test.go:31-152  ThrottleSpecBase.CalculateThreshold
# ------------------------------------------------------------------------------------------------
NOW = "2006-01-02T15:04:05Z"
SPEC_THRESHOLD = {"resourceCounts": {"pod": 0}, "resourceRequests": {"cpu": "1"}}
OVERRIDE1 = {"begin": "2006-01-02T15:03:05Z", "end": "2006-01-02T15:05:05Z",
             "threshold": {"resourceCounts": {"pod": 2}, "resourceRequests": {"cpu": "2"}}}
OVERRIDE2 = {"begin": "2006-01-02T15:03:05Z", "end": "2006-01-02T15:05:05Z",
             "threshold": {"resourceCounts": {"pod": 3}, "resourceRequests": {"cpu": "3", "memory": "3"}}}
ERRORED = {"begin": "error", "end": "error"}


def calc(oracle, overrides):
    thr = {"kind": "Throttle", "metadata": {"name": "t", "namespace": "default"},
           "spec": {"throttlerName": "dummy", "threshold": SPEC_THRESHOLD, "temporaryThresholdOverrides": overrides}}
    return oracle.call("CalculateThreshold", throttle=thr, now=NOW)


def norm_amount(a):
    return (a.get("resourceCounts"), rl_values(a.get("resourceRequests", {})))


def test_calculate_threshold(oracle):
    r = calc(oracle, [])
    assert norm_amount(r["threshold"]) == norm_amount(SPEC_THRESHOLD) and "messages" not in r and r["calculatedAtSet"]
    r = calc(oracle, [OVERRIDE1])
    assert norm_amount(r["threshold"]) == norm_amount(OVERRIDE1["threshold"])
    r = calc(oracle, [OVERRIDE1, OVERRIDE2])  # merged: first active wins per count / per resource (Q7: spec is REPLACED)
    assert norm_amount(r["threshold"]) == ({"pod": 2}, {"cpu": 2, "memory": 3})
    r = calc(oracle, [OVERRIDE1, ERRORED])
    assert norm_amount(r["threshold"]) == norm_amount(OVERRIDE1["threshold"])
    assert r["messages"] == [
        'index 1: Failed to parse Begin: parsing time "error" as "2006-01-02T15:04:05Z07:00": cannot parse "error" as "2006"']


# temporary_threshold_override_test.go:40-101  IsActive
BEGIN, END = "2021-08-04T10:00:00Z", "2021-08-05T10:00:00Z"


def active(oracle, ovr, now, off=0):
    return oracle.call("TemporaryThresholdOverride.IsActive", override=ovr, now=now, nowOffsetSec=off)


def test_is_active(oracle):
    assert active(oracle, {}, "")["active"] and active(oracle, {}, BEGIN)["active"] and active(oracle, {}, END)["active"]
    for off, want in ((-1, False), (0, True), (1, True), (65535 * 3600, True)):
        assert active(oracle, {"begin": BEGIN}, BEGIN, off)["active"] is want
    for off, want in ((-65535 * 3600, True), (-1, True), (0, True), (1, False)):
        assert active(oracle, {"end": END}, END, off)["active"] is want
    both = {"begin": BEGIN, "end": END}
    for now, off, want in ((BEGIN, -1, False), (BEGIN, 0, True), (BEGIN, 1, True), (END, -1, True), (END, 0, True), (END, 1, False)):
        assert active(oracle, both, now, off)["active"] is want
    assert "error" in active(oracle, {"begin": "not-time"}, "")
    assert "error" in active(oracle, {"end": "not-time"}, "")


def test_rfc3339_offsets(oracle):
    a = oracle.call("ParseRFC3339", value="2019-02-01T00:00:00+09:00")
    b = oracle.call("ParseRFC3339", value="2019-01-31T15:00:00Z")
    assert a["unix"] == b["unix"] == 1548946800
    assert oracle.call("ParseRFC3339", value="2021-08-04T10:00:00.5Z")["nsec"] == 500000000
    assert "month out of range" in oracle.call("ParseRFC3339", value="2021-13-04T10:00:00Z")["error"]
    assert "day out of range" in oracle.call("ParseRFC3339", value="2021-02-30T10:00:00Z")["error"]


# ------------------------------------------------------------------------------------------------
# throttle_selector_test.go:29-102, clusterthrottle_selector_test.go:29-110
# ------------------------------------------------------------------------------------------------
def test_throttle_selector(oracle):
    m = lambda sel, pod: oracle.call("ThrottleSelector.MatchesToPod", selector=sel, pod=pod)["match"]
    assert m({}, mk_pod(labels={"test": "test"})) is False  # no terms: matches nothing
    sel = {"selectorTerms": [{"podSelector": {"matchLabels": {"test1": "test1"}}}, {"podSelector": {"matchLabels": {"test2": "test2"}}}]}
    assert m(sel, mk_pod("test1", "test1", {"test1": "test1"})) is True
    assert m(sel, mk_pod("test2", "test2", {"test2": "test2"})) is True
    assert m(sel, mk_pod("test1", "test2", {"test1": "test2"})) is False
    empty_term = {"selectorTerms": [{}]}  # empty term: Everything
    assert m(empty_term, mk_pod(labels={"test": "test"})) is True
    assert m(empty_term, mk_pod()) is True


def test_clusterthrottle_selector(oracle):
    m = lambda sel, pod, ns: oracle.call("ClusterThrottleSelector.MatchesToPod", selector=sel, pod=pod, namespace=ns)["match"]
    assert m({}, mk_pod(labels={"test": "test"}), mk_ns("test", {"test": "test"})) is False
    l1, l2 = {"test1": "test1"}, {"test2": "test2"}
    sel = {"selectorTerms": [{"namespaceSelector": {"matchLabels": l1}, "podSelector": {"matchLabels": l1}},
                             {"namespaceSelector": {"matchLabels": l2}, "podSelector": {"matchLabels": l2}}]}
    assert m(sel, mk_pod("test1", "test1", l1), mk_ns("test1", l1)) is True
    assert m(sel, mk_pod("test2", "test2", l2), mk_ns("test2", l2)) is True
    assert m(sel, mk_pod("test1", "test2", l2), mk_ns("test1", l1)) is False  # ns and pod must match the SAME term
    empty_term = {"selectorTerms": [{}]}
    assert m(empty_term, mk_pod(labels={"test": "test"}), mk_ns("test1", {"test": "test"})) is True
    assert m(empty_term, mk_pod(), mk_ns("test1", None)) is True


# ------------------------------------------------------------------------------------------------
# resourcelist_test.go:47-421
# ------------------------------------------------------------------------------------------------
def test_pod_request_resource_list(oracle):
    pod = {"kind": "Pod", "metadata": {"name": "p"}, "spec": {"containers": [
        {"resources": {"requests": {"n1": "1"}}}, {"resources": {"requests": {"n1": "1"}}}]}}
    assert rl_values(oracle.call("PodRequestResourceList", pod=pod)) == {"n1": 2}
    pod["spec"]["initContainers"] = [{"resources": {"requests": {"n1": "1"}}}, {"resources": {"requests": {"n2": "2"}}}]
    assert rl_values(oracle.call("PodRequestResourceList", pod=pod)) == {"n1": 2, "n2": 2}
    pod["spec"]["overhead"] = {"n1": "500m", "n3": "1"}
    assert rl_values(oracle.call("PodRequestResourceList", pod=pod)) == {"n1": Fraction(5, 2), "n2": 2, "n3": 1}


@pytest.mark.parametrize("fn,lhs,rhs,want", [
    ("Add", {"n1": "0", "n2": "1", "n3": "1"}, {"n2": "0", "n3": "1", "n4": "2"}, {"n1": 0, "n2": 1, "n3": 2, "n4": 2}),
    ("Add", {"n1": "0", "n2": "0", "n3": "1", "n4": "1"}, {"n1": "0", "n2": "1", "n3": "0", "n4": "1"}, {"n1": 0, "n2": 1, "n3": 1, "n4": 2}),
    ("Sub", {"n1": "1", "n2": "1", "n3": "1"}, {"n2": "0", "n3": "1", "n4": "2"}, {"n1": 1, "n2": 1, "n3": 0, "n4": -2}),
    ("Sub", {"n1": "0", "n2": "0", "n3": "1", "n4": "1"}, {"n1": "0", "n2": "1", "n3": "0", "n4": "1"}, {"n1": 0, "n2": -1, "n3": 1, "n4": 0}),
    ("SetMax", {"n1": "1", "n2": "2", "n3": "2"}, {"n2": "1", "n3": "2", "n4": "0"}, {"n1": 1, "n2": 2, "n3": 2, "n4": 0}),
    ("SetMax", {"n1": "1", "n2": "1", "n3": "2", "n4": "2"}, {"n1": "1", "n2": "2", "n3": "1", "n4": "2"}, {"n1": 1, "n2": 2, "n3": 2, "n4": 2}),
    ("SetMin", {"n1": "1", "n2": "2", "n3": "1"}, {"n2": "1", "n3": "2", "n4": "1"}, {"n2": 1, "n3": 1}),
    ("SetMin", {"n1": "1", "n2": "2", "n3": "0", "n4": "2"}, {"n1": "2", "n2": "0", "n3": "2", "n4": "1"}, {"n1": 1, "n2": 0, "n3": 0, "n4": 1}),
])
def test_resourcelist_algebra(oracle, fn, lhs, rhs, want):
    assert rl_values(oracle.call("ResourceList." + fn, lhs=lhs, rhs=rhs)) == want


@pytest.mark.parametrize("lhs,rhs,want", [
    ({"n1": "1", "n2": "2", "n3": "2"}, {"n1": "1", "n2": "1", "n3": "2"}, True),
    ({"n1": "1", "n2": "2", "n3": "2"}, {"n1": "1", "n3": "2"}, True),
    ({"n1": "1", "n2": "1", "n3": "2"}, {"n1": "1", "n2": "2", "n3": "1"}, False),
    ({"n1": "1", "n2": "2", "n3": "1"}, {"n1": "1", "n3": "2"}, False),
    ({"n1": "1", "n2": "2", "n3": "1"}, {"n1": "1", "n2": "2", "n3": "1"}, True),
    ({"n1": "1", "n2": "2", "n3": "1"}, {"n1": "1", "n3": "1"}, True),
    ({"n1": "1"}, {"n1": "1", "n2": "0"}, False),  # missing key in lhs => false even against zero
])
def test_resourcelist_greater_or_equal(oracle, lhs, rhs, want):
    assert oracle.call("ResourceList.GreaterOrEqual", lhs=lhs, rhs=rhs) is want


def test_resourcelist_equal_to(oracle):
    assert oracle.call("ResourceList.EqualTo", lhs={"n1": "0"}, rhs={}) is True  # missing compares as zero
    assert oracle.call("ResourceList.EqualTo", lhs={"n1": "1"}, rhs={}) is False
    assert oracle.call("ResourceList.EqualTo", lhs={"n1": "1000m"}, rhs={"n1": "1"}) is True


# ------------------------------------------------------------------------------------------------
# resource.Quantity: apimachinery v0.26.4 semantics; beyond plain ints and `m` the reference's tests
# do not pin these (README.md:287-309 shows 512Mi == 536870912 once): PARITY UNPINNED, spec-derived.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("s,want", [
    ("0", 0), ("1", 1), ("-2", -2), ("+3", 3), ("500m", Fraction(1, 2)), ("1.1", Fraction(11, 10)), ("0.200", Fraction(1, 5)),
    ("512Mi", 536870912), ("1Gi", 2**30), ("1Ki", 1024), ("1.5Gi", 3 * 2**29), ("100n", Fraction(1, 10**7)), ("5u", Fraction(5, 10**6)),
    ("3k", 3000), ("2M", 2 * 10**6), ("1G", 10**9), ("1T", 10**12), ("1P", 10**15), ("1E", 10**18), ("1e3", 1000), ("1E-3", Fraction(1, 1000)),
    ("12e6", 12 * 10**6), (".5", Fraction(1, 2)), ("5.", 5),
    ("0.0000000001", Fraction(1, 10**9)),      # finer than nano: rounded UP to 1n
    ("1.0000000001", Fraction(10**9 + 1, 10**9)),
    ("-0.0000000001", Fraction(-1, 10**9)),    # magnitude rounded up, sign kept
    ("16Ei", 2**63 - 1),                       # BinarySI cap at MaxInt64
])
def test_parse_quantity(oracle, s, want):
    assert Fraction(oracle.call("ParseQuantity", value=s)["decimal"]) == want


@pytest.mark.parametrize("s", ["", "abc", "1x", "1Kii", "--1", "1e", "1e1.5", "..", "1.5.5", "1,5"])
def test_parse_quantity_errors(oracle, s):
    with pytest.raises(RuntimeError):
        oracle.call("ParseQuantity", value=s)


def test_quantity_cmp_across_scales(oracle):
    assert oracle.call("Quantity.Cmp", a="1000m", b="1") == 0
    assert oracle.call("Quantity.Cmp", a="1Gi", b="1G") == 1
    assert oracle.call("Quantity.Cmp", a="999m", b="1") == -1


# ------------------------------------------------------------------------------------------------
# matchExpressions (apimachinery semantics; PARITY UNPINNED by the reference's tests)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("expr,labels,want", [
    ({"key": "a", "operator": "In", "values": ["x", "y"]}, {"a": "x"}, True),
    ({"key": "a", "operator": "In", "values": ["x", "y"]}, {"a": "z"}, False),
    ({"key": "a", "operator": "In", "values": ["x"]}, {}, False),
    ({"key": "a", "operator": "NotIn", "values": ["x"]}, {}, True),      # absent key passes NotIn
    ({"key": "a", "operator": "NotIn", "values": ["x"]}, {"a": "x"}, False),
    ({"key": "a", "operator": "NotIn", "values": ["x"]}, {"a": "y"}, True),
    ({"key": "a", "operator": "Exists"}, {"a": ""}, True),
    ({"key": "a", "operator": "Exists"}, {"b": "1"}, False),
    ({"key": "a", "operator": "DoesNotExist"}, {"b": "1"}, True),
    ({"key": "a", "operator": "DoesNotExist"}, {"a": "1"}, False),
])
def test_match_expressions(oracle, expr, labels, want):
    sel = {"selectorTerms": [{"podSelector": {"matchExpressions": [expr]}}]}
    assert oracle.call("ThrottleSelector.MatchesToPod", selector=sel, pod=mk_pod(labels=labels))["match"] is want


@pytest.mark.parametrize("expr", [
    {"key": "a", "operator": "In", "values": []},
    {"key": "a", "operator": "Exists", "values": ["x"]},
    {"key": "a", "operator": "Bogus"},
    {"key": "not a key!", "operator": "Exists"},
    {"key": "a", "operator": "In", "values": ["bad value!"]},
])
def test_selector_errors(oracle, expr):
    sel = {"selectorTerms": [{"podSelector": {"matchExpressions": [expr]}}]}
    r = oracle.call("ThrottleSelector.MatchesToPod", selector=sel, pod=mk_pod(labels={"a": "x"}))
    assert "error" in r and r["match"] is False
    # the same broken selector on the NAMESPACE side is swallowed (Q9, clusterthrottle_selector.go:63-77)
    csel = {"selectorTerms": [{"namespaceSelector": {"matchExpressions": [expr]}}]}
    r = oracle.call("ClusterThrottleSelector.MatchesToPod", selector=csel, pod=mk_pod(labels={"a": "x"}), namespace=mk_ns("n", {"a": "x"}))
    assert r == {"match": False}
