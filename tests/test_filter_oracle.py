"""The hybrid-filter oracle (oracle/filter_oracle.py, a plain-Python restatement of the reference's TAG / NUMERIC /
predicate semantics) is (1) pinned to RediSearch's recorded answers and to the reference's own unit-test expectations,
then (2) used as the checker of the C++ host mirror (tests/native/reference_filter_standins.cc) on random inputs: the
same records, mutations and predicate trees go through both, every status and every selected key set must agree."""
import json
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import filter_oracle as F  # noqa: E402

BIN = os.path.join(ROOT, "tests", "native", "filter_index_test")
hx = lambda s: s.encode("utf-8").hex() or "-"  # "-" stands for the empty string


# ---------------------------------------------------------------------------------------------- pinning the oracle
def test_oracle_reproduces_redisearch_tag_answers():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "redisearch_tag_special_chars.json")))
    ix = F.TagIndex(g["separator"], g["case_sensitive"])
    for k, v in g["docs"]:
        assert ix.add(k, v) == "added"
    for c in g["cases"]:
        q = c["query"]
        tag_string = F.parse_tag_string(q[q.index("{") + 1:])
        keys = F.prefiltered_keys(("tag", "t", tag_string), [k for k, _ in g["docs"]], {"t": ix})
        assert keys == c["keys"] and len(keys) == c["count"], q


def test_oracle_tag_cases_of_the_reference_unit_tests():
    """testing/tag_index_test.cc:52-135 (add / remove / modify / tracking), :137-207 (prefix), :281-435 (escapes)."""
    ix = F.TagIndex(",", False)
    assert ix.add("key1", "    ") == "missing" and ix.add("key1", "tag1") == "added" and ix.add("key2", "tag2") == "added"
    with pytest.raises(F.FilterError):
        ix.add("key2", "tag2")
    assert F.prefiltered_keys(("tag", "t", "tag1"), ["key1", "key2"], {"t": ix}) == ["key1"]
    assert ix.modify("key1", "tag2.1,tag2.2") == "added"
    assert F.prefiltered_keys(("tag", "t", "tag2.1"), ["key1", "key2"], {"t": ix}) == ["key1"]
    with pytest.raises(F.FilterError):
        ix.modify("key5", "tag5")
    assert ix.modify("key1", "") == "missing" and "key1" not in ix.tracked and "key1" in ix.untracked
    assert ix.remove("key2") is True and ix.remove("key2") is False
    p = F.TagIndex(",", False)
    for k, v in (("doc1", "disagree"), ("doc2", "disappear"), ("doc3", "dislike"), ("doc4", "disadvantage"), ("doc5", "preschool")):
        p.add(k, v)
    for q in ("dis*", "dIs*"):
        assert F.prefiltered_keys(("tag", "t", q), [f"doc{i}" for i in range(1, 6)], {"t": p}) == ["doc1", "doc2", "doc3", "doc4"]
    for bad in ("dis**", "d*", "b*", "*"):
        with pytest.raises(F.FilterError):
            F.parse_search_tags(bad, "|")
    un = lambda raw: {F.unescape_tag(t) for t in F.parse_search_tags(raw, "|")}
    assert un(r"foo\|bar") == {"foo|bar"} and un(r"a\|b|c") == {"a|b", "c"} and un(r"foo\\|bar") == {"foo\\", "bar"}
    assert un(r"foo\\\|bar") == {r"foo\|bar"} and un(r"a\|b\|c|d\|e") == {"a|b|c", "d|e"}
    assert un(r"foo\\") == {"foo\\"} and un(r"foo\|") == {"foo|"} and un("a||b") == {"a", "b"} and un("a|   |b") == {"a", "b"}
    assert F.unescape_tag("abc\\") == "abc\\" and F.unescape_tag(r"a\|b\\c") == r"a|b\c" and F.unescape_tag(r"test\value") == "testvalue"
    assert F.parse_search_tags("tag\\", "|") == {"tag\\"} and F.parse_search_tags("", "|") == set() == F.parse_search_tags("   ", "|")


def test_oracle_numeric_cases_of_the_reference_unit_tests():
    """testing/numeric_index_test.cc:43-108, 147-191."""
    ix = F.NumericIndex()
    assert ix.add("key1", "not_a_number") == "invalid" and ix.add("key2", "nan") == "invalid" and ix.add("key3", "") == "invalid"
    assert ix.add("key4", "42") == "added" and ix.modify("key4", "still_not_a_number") == "invalid" and "key4" not in ix.tracked
    ix = F.NumericIndex()
    for i, v in enumerate(("1.0", "2.0", "2.2", "3.2", "2.0", "2.1"), 1):
        assert ix.add(f"key{i}", v) == "added"
    keys = [f"key{i}" for i in range(1, 7)]
    sel = lambda a, ia, b, ib: F.prefiltered_keys(("num", "n", a, ia, b, ib), keys, {"n": ix})
    assert sel(1.0, True, 2.1, True) == ["key1", "key2", "key5", "key6"]
    assert sel(1.0, False, 2.1, True) == ["key2", "key5", "key6"]
    assert sel(1.0, False, 2.1, False) == ["key2", "key5"]
    assert sel(1.0, False, 3.5, False) == ["key2", "key3", "key4", "key5", "key6"]
    assert sel(0.0, False, 2.1, False) == ["key1", "key2", "key5"]


# ---------------------------------------------------------------------------------------------- differential test
TAG_ALPHABET = ["red", "RED", "Red", "green", "blue", "dark blue", "darkness", "dis", "disagree", "dislike", "a|b", "a}b",
                "a\\b", "x\\", "café", "中文", "😀", " padded ", "tab\there", "q*", "star*x", ""]
NUMBERS = ["0", "-0", "1", "1.5", "-2.25", "1e3", "-1E-2", " 7 ", "inf", "-inf", "+3", ".5", "5.", "bad", "nan", "NaN", "",
           "0x10", "1.5abc", "12 3", "1e999", "+-3", "-+3", "+ 3", " nan", "-nan", "Infinity", "-INF", "1e", "e5", ".", "-.5e1",
           "1_000", "-0x1p3", "1e-999", "00012", "1.e2"]


CHARSET = ["a", "b", "A", "B", "\\", "|", "*", " ", ",", ";", "}", "\t", "é", "x"]


def _random_text(rng, lo=0, hi=7):
    return "".join(rng.choice(CHARSET) for _ in range(rng.randint(lo, hi)))


def _random_tree(rng, depth=0):
    r = rng.random()
    if depth >= 3 or r < 0.35:
        if rng.random() < 0.6:
            pieces = []
            for _ in range(rng.randint(1, 3)):
                t = rng.choice(["red", "RED", "green", "blu*", "dark*", "dis*", "disagree", "a\\|b", "a\\}b", "a\\\\b", "x\\\\",
                                "café", "中文", "😀", "padded", "tab\\\there", "q\\*", "nomatch", "  ", "da*"])
                pieces.append(t)
            if rng.random() < 0.35:  # arbitrary text: escapes, separators, wildcards and blanks in any position
                return ("tag", rng.choice(["t1", "t2"]), _random_text(rng, 0, 9))
            return ("tag", rng.choice(["t1", "t2"]), " | ".join(pieces))
        a, b = sorted(rng.choice([-3.0, -0.01, 0.0, 0.5, 1.0, 1.5, 3.0, 7.0, 1000.0, float("-inf"), float("inf")]) for _ in range(2))
        return ("num", rng.choice(["n1", "n2"]), a, rng.random() < 0.5, b, rng.random() < 0.5)
    if r < 0.5:
        return ("not", _random_tree(rng, depth + 1))
    return ("and" if r < 0.75 else "or", [_random_tree(rng, depth + 1) for _ in range(rng.randint(1, 3))])


def _tokens(tree):
    k = tree[0]
    if k == "tag":
        return ["TAG", tree[1], hx(tree[2])]
    if k == "num":
        return ["NUM", tree[1], F.fmt_double(tree[2]), "1" if tree[3] else "0", F.fmt_double(tree[4]), "1" if tree[5] else "0"]
    if k == "not":
        return ["NOT"] + _tokens(tree[1])
    out = [k.upper(), str(len(tree[1]))]
    for c in tree[1]:
        out += _tokens(c)
    return out


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_cpp_host_mirror_agrees_with_the_oracle_on_random_inputs(built, tmp_path, seed):
    rng = random.Random(seed)
    sep = rng.choice([",", ";", "|"])
    tags = {"t1": F.TagIndex(sep, False), "t2": F.TagIndex(sep, True)}
    nums = {"n1": F.NumericIndex(), "n2": F.NumericIndex()}
    lines, expect = [], []
    for name, ix in tags.items():
        lines.append(f"tagindex {name} {hx(sep)} {1 if ix.case_sensitive else 0}")
    for name in nums:
        lines.append(f"numindex {name}")
    keys = [f"k{i}" for i in range(60)]

    def record(fn):
        try:
            expect.append("rec " + fn())
        except F.FilterError as e:
            expect.append("rec ERR " + str(e))

    for _ in range(400):
        key = rng.choice(keys)
        op = rng.random()
        if rng.random() < 0.55:
            name = rng.choice(list(tags))
            data = sep.join(rng.choice(TAG_ALPHABET) for _ in range(rng.randint(0, 3)))
            if rng.random() < 0.3:
                data = _random_text(rng, 0, 9)
            if op < 0.5:
                lines.append(f"tadd {name} {hx(key)} {hx(data)}")
                record(lambda: tags[name].add(key, data))
            elif op < 0.8:
                lines.append(f"tmod {name} {hx(key)} {hx(data)}")
                record(lambda: tags[name].modify(key, data))
            else:
                how = rng.choice(["none", "record"])
                lines.append(f"trem {name} {hx(key)} {how}")
                expect.append(f"rem {1 if tags[name].remove(key, how) else 0}")
        else:
            name = rng.choice(list(nums))
            data = rng.choice(NUMBERS)
            if op < 0.5:
                lines.append(f"nadd {name} {hx(key)} {hx(data)}")
                record(lambda: nums[name].add(key, data))
            elif op < 0.8:
                lines.append(f"nmod {name} {hx(key)} {hx(data)}")
                record(lambda: nums[name].modify(key, data))
            else:
                how = rng.choice(["none", "record"])
                lines.append(f"nrem {name} {hx(key)} {how}")
                expect.append(f"rem {1 if nums[name].remove(key, how) else 0}")
    universe = keys[:50] + ["never-seen"]
    lines += [f"universe {hx(k)}" for k in universe]
    indexes = {**tags, **nums}
    for _ in range(150):
        tree = _random_tree(rng)
        lines.append("pred " + " ".join(_tokens(tree)))
        try:
            sel = F.prefiltered_keys(tree, universe, indexes)
            expect.append("pred " + " ".join([str(len(sel))] + [hx(k) for k in sorted(sel)]))
        except F.FilterError as e:
            expect.append("pred ERR " + str(e))
    path = tmp_path / "eval.txt"
    path.write_text("\n".join(lines) + "\n")
    p = subprocess.run([BIN, "--eval", str(path)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    got = p.stdout.strip().splitlines()
    assert len(got) == len(expect)
    for i, (g, e) in enumerate(zip(got, expect)):
        assert g == e, (i, g, e)
