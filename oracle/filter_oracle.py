"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement, in plain Python, of the reference's hybrid-filter semantics (the "next" row N1 of SURVEY §8f): what
a TAG / NUMERIC attribute index accepts, how a query's tag clause is split and unescaped, and when a key satisfies a
predicate tree.  It is the checker for tests/native/reference_filter_standins.{h,cc} + valkey_search_b200/host/device_filter.{h,cc} (tests/test_filter_oracle.py runs
both on random inputs and compares) and is itself pinned to
  * RediSearch's recorded answers for the reference's `tag special chars` compatibility data set
    (tests/golden/redisearch_tag_special_chars.json, written by tests/golden/make_golden.py), and
  * the expectations of the reference's own unit tests (testing/tag_index_test.cc, testing/numeric_index_test.cc),
    re-stated in tests/test_filter_oracle.py.
Only tests/ may import this module.  Every function cites the reference lines it follows.
"""
import math

ASCII_WS = " \t\n\v\f\r"


def strip_ascii_whitespace(s):
    """absl::StripAsciiWhitespace."""
    return s.strip(ASCII_WS)


def ascii_lower(s):
    """absl::ascii_tolower per byte: only A-Z change (tag.cc:81-90)."""
    return "".join(chr(ord(c) + 32) if "A" <= c <= "Z" else c for c in s)


# ---------------------------------------------------------------------------------------------- TAG
def parse_record_tags(data, separator):
    """Tag::ParseRecordTags, src/indexes/tag.cc:196-206: split at every separator, strip, drop empties."""
    return {t for t in (strip_ascii_whitespace(p) for p in data.split(separator)) if t}


def is_valid_prefix(s):
    """tag.cc:66-69."""
    return len(s) < 2 or s[-1] != "*" or s[-2] != "*"


class FilterError(Exception):
    pass


def parse_search_tags(data, separator, min_prefix_length=2):
    """Tag::ParseSearchTags, tag.cc:145-194: a backslash escapes the next character (so an escaped separator does not
    split); pieces are stripped, empties ignored; a trailing '*' makes a prefix query that must not end in '**' and
    must be longer than tag-min-prefix-length (counting the '*')."""
    out = set()

    def insert(raw):
        tag = strip_ascii_whitespace(raw)
        if not tag:
            return
        if tag[-1] == "*":
            if not is_valid_prefix(tag):
                raise FilterError(f"Tag string `{tag}` ends with multiple *.")
            if len(tag.encode("utf-8")) <= min_prefix_length:  # the reference measures bytes
                raise FilterError(f"Tag string `{tag}` is too short for prefix wildcard.")
        out.add(tag)

    start, i = 0, 0
    while i < len(data):
        if data[i] == "\\" and i + 1 < len(data):
            i += 1
        elif data[i] == separator:
            insert(data[start:i])
            start = i + 1
        i += 1
    insert(data[start:])
    return out


def unescape_tag(tag):
    """Tag::UnescapeTag, tag.cc:131-143: backslash + c -> c; a trailing lone backslash stays."""
    out, i = [], 0
    while i < len(tag):
        if tag[i] == "\\" and i + 1 < len(tag):
            i += 1
        out.append(tag[i])
        i += 1
    return "".join(out)


def parse_tag_string(expression):
    """FilterParser::ParseTagString, src/commands/filter_parser.cc:329-348: text up to the first unescaped '}'."""
    i = 0
    while i < len(expression):
        if expression[i] == "\\" and i + 1 < len(expression):
            i += 1
        elif expression[i] == "}":
            return expression[:i]
        i += 1
    raise FilterError("Missing closing TAG bracket, '}'")


def query_tags(tag_string):
    """FilterParser::ParseQueryTags (filter_parser.cc:350-357: '|' always separates query tags) followed by the
    unescaping TagPredicate's constructor applies (src/query/predicate.cc:343-356)."""
    return {unescape_tag(t) for t in parse_search_tags(tag_string, "|")}


def tag_predicate_evaluate(in_tags, tags, case_sensitive):
    """TagPredicate::Evaluate, predicate.cc:362-393.  in_tags: the key's parsed tag set or None."""
    if in_tags is None:
        return False
    for in_tag in in_tags:
        for tag in tags:
            lhs, rhs = in_tag.encode("utf-8"), tag.encode("utf-8")  # the reference compares bytes
            if rhs and rhs[-1:] == b"*":
                if len(lhs) < len(rhs) - 1:
                    continue
                lhs = lhs[: len(rhs) - 1]
                rhs = rhs[:-1]
            if (lhs == rhs) if case_sensitive else (_lower_bytes(lhs) == _lower_bytes(rhs)):
                return True
    return False


def _lower_bytes(b):
    return bytes(c + 32 if 65 <= c <= 90 else c for c in b)


class TagIndex:
    """indexes::Tag's record side (tag.cc:107-129, 208-265): key -> raw tag string for keys with at least one tag;
    keys seen without one are `untracked`."""

    def __init__(self, separator=",", case_sensitive=False):
        self.separator, self.case_sensitive = separator, case_sensitive
        self.tracked, self.untracked = {}, set()

    def add(self, key, data):
        if not parse_record_tags(data, self.separator):
            self.untracked.add(key)
            return "missing"
        if key in self.tracked:
            raise FilterError(f"Key `{key}` already exists")
        self.tracked[key] = data
        self.untracked.discard(key)
        return "added"

    def modify(self, key, data):
        if not parse_record_tags(data, self.separator):
            self.remove(key, "identifier")
            return "missing"
        if key not in self.tracked:
            raise FilterError(f"Key `{key}` not found")
        self.tracked[key] = data
        return "added"

    def remove(self, key, deletion_type="none"):
        if deletion_type == "record":
            self.untracked.discard(key)
        else:
            self.untracked.add(key)
        return self.tracked.pop(key, None) is not None

    def value(self, key):
        data = self.tracked.get(key)
        return None if data is None else parse_record_tags(data, self.separator)


# ---------------------------------------------------------------------------------------------- NUMERIC
def parse_number(data):
    """ParseNumber, src/indexes/numeric.cc:30-36: the exact text "nan" (any case, nothing around it) is refused;
    everything else goes through absl::SimpleAtod — surrounding whitespace allowed, one leading '+' allowed unless a
    '-' follows, decimal or scientific notation, inf / infinity / nan spellings, no hexadecimal, no digit separators,
    the whole text must be consumed, overflow gives +-infinity.  So " nan" or "-nan" is accepted, as a NaN."""
    import re
    if data.lower() == "nan":
        return None
    s = strip_ascii_whitespace(data)
    if s.startswith("+"):
        s = s[1:]
        if s.startswith("-"):
            return None
    if not s or s[0] in ASCII_WS:
        return None
    if not re.fullmatch(r"-?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|inf|infinity|nan(\([A-Za-z0-9_]*\))?)", s, re.I | re.A):
        return None
    if "nan" in s.lower():
        return float("nan")
    return float(s)


def numeric_predicate_evaluate(value, start, inclusive_start, end, inclusive_end):
    """NumericPredicate::Evaluate, predicate.cc:332-341."""
    if value is None:
        return False
    return ((value > start or (inclusive_start and value == start)) and value < end) or (inclusive_end and value == end)


class NumericIndex:
    """indexes::Numeric's record side (numeric.cc:44-106)."""

    def __init__(self):
        self.tracked, self.untracked = {}, set()

    def add(self, key, data):
        v = parse_number(data)
        if v is None:
            self.untracked.add(key)
            return "invalid"
        if key in self.tracked:
            raise FilterError(f"Key `{key}` already exists")
        self.tracked[key] = v
        self.untracked.discard(key)
        return "added"

    def modify(self, key, data):
        v = parse_number(data)
        if v is None:
            self.remove(key, "identifier")
            return "invalid"
        if key not in self.tracked:
            raise FilterError(f"Key `{key}` not found")
        self.tracked[key] = v
        return "added"

    def remove(self, key, deletion_type="none"):
        if deletion_type == "record":
            self.untracked.discard(key)
        else:
            self.untracked.add(key)
        return self.tracked.pop(key, None) is not None


# ---------------------------------------------------------------------------------------------- predicate trees
def evaluate(tree, key, indexes):
    """A predicate tree on one key (predicate.cc:36-39 NOT, 429-520 AND / OR without text children).  Trees are
    nested tuples: ("tag", index_name, tag_string) | ("num", index_name, start, incl_start, end, incl_end) |
    ("and", [children]) | ("or", [children]) | ("not", child)."""
    kind = tree[0]
    if kind == "tag":
        ix = indexes[tree[1]]
        return tag_predicate_evaluate(ix.value(key), query_tags(tree[2]), ix.case_sensitive)
    if kind == "num":
        return numeric_predicate_evaluate(indexes[tree[1]].tracked.get(key), tree[2], tree[3], tree[4], tree[5])
    if kind == "not":
        return not evaluate(tree[1], key, indexes)
    if kind == "and":
        return all(evaluate(c, key, indexes) for c in tree[1])
    if kind == "or":
        return any(evaluate(c, key, indexes) for c in tree[1])
    raise ValueError(kind)


def validate(tree):
    """The query is parsed before anything is evaluated (FilterParser::Parse builds every TagPredicate up front,
    src/commands/filter_parser.cc:359-376), so a malformed tag clause fails the query even where evaluation would
    have short-circuited past it."""
    kind = tree[0]
    if kind == "tag":
        query_tags(tree[2])
    elif kind == "not":
        validate(tree[1])
    elif kind in ("and", "or"):
        for c in tree[1]:
            validate(c)


def prefiltered_keys(tree, universe, indexes):
    """EvaluatePrefilteredKeys (src/query/search.cc:401-455) reduced to its result: the keys of the vector index
    (`universe`) for which the root predicate holds."""
    validate(tree)
    return sorted(k for k in universe if evaluate(tree, k, indexes))


def fmt_double(x):
    """A double as text the C++ side parses back to the same bits."""
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return repr(float(x))
