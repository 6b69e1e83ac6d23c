// Tests of the TAG / NUMERIC candidate-set bridge (valkey_search_b200/host/device_filter.h driven by tests/native/reference_filter_standins.h).
//   host cases  = the reference's testing/tag_index_test.cc and testing/numeric_index_test.cc re-stated (same inputs,
//                 same expectations) + the predicate tree's Evaluate();  run everywhere (`--host-only`)
//   device case = on a B200: for a corpus with tags and prices attached in every order, mutated, for a list of
//                 predicate trees — the label set computed ON THE DEVICE (posting bitmaps, range kernel, set
//                 algebra) equals the set the reference's per-key evaluation gives, and the filtered kNN through it
//                 returns exactly what the already-verified key-list pre-filter path returns
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <type_traits>
#include <vector>

#include "reference_filter_standins.h"

using namespace valkey_search::indexes;

static int g_failures = 0, g_checks = 0;
#define EXPECT_TRUE(c)                                                      \
  do {                                                                      \
    g_checks++;                                                             \
    if (!(c)) {                                                             \
      g_failures++;                                                         \
      std::fprintf(stderr, "  FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
    }                                                                       \
  } while (0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_OK(s) EXPECT_TRUE((s).ok())

using Keys = std::multiset<std::string>;
static Keys AsSet(const std::vector<std::string> &v) { return Keys(v.begin(), v.end()); }
static std::set<std::string> Unescaped(const std::string &raw, char sep) {  // ParseAndUnescapeTags, tag_index_test.cc:268-278
  auto parsed = Tag::ParseSearchTags(raw, sep);
  std::set<std::string> out;
  if (!parsed.ok()) return out;
  for (const auto &t : *parsed) out.insert(Tag::UnescapeTag(t));
  return out;
}
static TagPredicate QueryTags(Tag *index, const std::string &filter) {  // FilterParser::ParseQueryTags: '|' separates
  return TagPredicate(index, *Tag::ParseSearchTags(filter, '|'));
}

// ------------------------------------------------------------------------------------------ tag_index_test.cc
static void TagIndexCases() {
  {  // AddRecordAndSearchTest :52-69
    Tag index(',', false);
    EXPECT_EQ(*index.AddRecord("key1", "    "), RecordResult::kMissing);
    EXPECT_EQ(*index.AddRecord("key1", "tag1"), RecordResult::kAdded);
    EXPECT_EQ(*index.AddRecord("key2", "tag2"), RecordResult::kAdded);
    EXPECT_EQ(index.AddRecord("key2", "tag2").status().code(), vks::StatusCode::kAlreadyExists);
    EXPECT_EQ(AsSet(index.Search(QueryTags(&index, "tag1"), false)), Keys({"key1"}));
  }
  {  // RemoveRecordAndSearchTest :71-85
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("key1", "tag1"));
    EXPECT_OK(index.AddRecord("key2", "tag2"));
    EXPECT_TRUE(*index.RemoveRecord("key1"));
    EXPECT_EQ(index.Search(QueryTags(&index, "tag1"), false).size(), (size_t)0);
  }
  {  // ModifyRecordAndSearchTest :87-103
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("key1", "tag2"));
    EXPECT_EQ(*index.ModifyRecord("key1", "tag2.1,tag2.2"), RecordResult::kAdded);
    EXPECT_EQ(AsSet(index.Search(QueryTags(&index, "tag2.1"), false)), Keys({"key1"}));
    EXPECT_EQ(index.Search(QueryTags(&index, "tag2"), false).size(), (size_t)0);
    EXPECT_EQ(index.ModifyRecord("key5", "tag5").status().code(), vks::StatusCode::kNotFound);
  }
  {  // ModifyRecordWithEmptyString :105-119
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("key1", "tag2"));
    EXPECT_EQ(*index.ModifyRecord("key1", ""), RecordResult::kMissing);
    EXPECT_EQ(index.Search(QueryTags(&index, "tag2"), false).size(), (size_t)0);
    EXPECT_EQ(index.GetTrackedKeyCount(), (size_t)0);
  }
  {  // KeyTrackingTest :121-135
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("key1", "tag1"));
    EXPECT_OK(index.AddRecord("key2", "tag2"));
    EXPECT_FALSE(index.IsTracked("key3"));
    EXPECT_OK(index.AddRecord("key3", "tag3"));
    EXPECT_TRUE(index.IsTracked("key3"));
    EXPECT_TRUE(*index.RemoveRecord("key3"));
    EXPECT_FALSE(index.IsTracked("key3"));
    auto res = index.RemoveRecord("key3");
    EXPECT_TRUE(res.ok() && !*res);
    EXPECT_EQ(*index.AddRecord("key5", "  "), RecordResult::kMissing);
    EXPECT_EQ(*index.ModifyRecord("key5", " "), RecordResult::kMissing);
    EXPECT_EQ(*index.AddRecord("key6", " tag6 , tag7 "), RecordResult::kAdded);
    EXPECT_EQ(*index.GetValue("key6"), std::set<std::string>({"tag6", "tag7"}));
  }
  for (const char *filter : {"dis*", "dIs*"}) {  // PrefixSearchHappyTest / CaseInsensitive :137-175
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("doc1", "disagree"));
    EXPECT_OK(index.AddRecord("doc2", "disappear"));
    EXPECT_OK(index.AddRecord("doc3", "dislike"));
    EXPECT_OK(index.AddRecord("doc4", "disadvantage"));
    EXPECT_OK(index.AddRecord("doc5", "preschool"));
    auto parsed = Tag::ParseSearchTags(filter, '|');
    EXPECT_OK(parsed);
    EXPECT_EQ(*parsed, std::set<std::string>({filter}));
    EXPECT_EQ(AsSet(index.Search(TagPredicate(&index, *parsed), false)), Keys({"doc1", "doc2", "doc3", "doc4"}));
  }
  // PrefixSearchInvalidTagTest / MinLength :177-207
  EXPECT_EQ(Tag::ParseSearchTags("dis**", ',').status().code(), vks::StatusCode::kInvalidArgument);
  EXPECT_EQ(Tag::ParseSearchTags("d*", '|').status().code(), vks::StatusCode::kInvalidArgument);
  EXPECT_EQ(Tag::ParseSearchTags("dis*", '|')->size(), (size_t)1);
  {  // NegativeSearchTest :209-235
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("doc1", "disagree"));
    EXPECT_OK(index.AddRecord("doc2", "distance"));
    EXPECT_TRUE(*index.RemoveRecord("doc1"));  // now untracked
    EXPECT_TRUE(*index.RemoveRecord("doc2"));
    EXPECT_OK(index.AddRecord("doc3", "decorum"));
    EXPECT_OK(index.AddRecord("doc4", "dismiss"));
    EXPECT_FALSE(*index.RemoveRecord("doc5"));  // removed, never added
    EXPECT_OK(index.AddRecord("doc6", "demand"));
    EXPECT_TRUE(*index.RemoveRecord("doc6"));
    EXPECT_OK(index.AddRecord("doc6", "demand2"));  // removed then added
    EXPECT_EQ(AsSet(index.Search(QueryTags(&index, "dis*"), true)), Keys({"doc1", "doc2", "doc3", "doc5", "doc6"}));
  }
  {  // DeletedKeysNegativeSearchTest :237-266
    Tag index(',', false);
    EXPECT_OK(index.AddRecord("doc0", "ambiance"));
    EXPECT_OK(index.AddRecord("doc1", "demand"));
    EXPECT_TRUE(*index.RemoveRecord("doc1", DeletionType::kNone));  // the field went away
    EXPECT_EQ(AsSet(index.Search(QueryTags(&index, "dis*"), true)), Keys({"doc0", "doc1"}));
    EXPECT_FALSE(*index.RemoveRecord("doc1", DeletionType::kRecord));  // the key went away
    EXPECT_EQ(AsSet(index.Search(QueryTags(&index, "dis*"), true)), Keys({"doc0"}));
  }
  // escapes :281-337
  EXPECT_EQ(Unescaped(R"(foo\|bar)", '|'), std::set<std::string>({"foo|bar"}));
  EXPECT_EQ(Unescaped(R"(a\|b|c)", '|'), std::set<std::string>({"a|b", "c"}));
  EXPECT_EQ(Unescaped(R"(foo\\|bar)", '|'), std::set<std::string>({R"(foo\)", "bar"}));
  EXPECT_EQ(Unescaped(R"(foo\\\|bar)", '|'), std::set<std::string>({R"(foo\|bar)"}));
  EXPECT_EQ(Unescaped(R"(a\|b\|c|d\|e)", '|'), std::set<std::string>({"a|b|c", "d|e"}));
  EXPECT_EQ(Unescaped(R"(foo\\)", '|'), std::set<std::string>({R"(foo\)"}));
  EXPECT_EQ(Unescaped(R"(foo\|)", '|'), std::set<std::string>({"foo|"}));
  // UnescapeTag :339-378
  EXPECT_EQ(Tag::UnescapeTag(""), std::string(""));
  EXPECT_EQ(Tag::UnescapeTag("hello world"), std::string("hello world"));
  EXPECT_EQ(Tag::UnescapeTag(R"(a\|b)"), std::string("a|b"));
  EXPECT_EQ(Tag::UnescapeTag(R"(a\\b)"), std::string(R"(a\b)"));
  EXPECT_EQ(Tag::UnescapeTag(R"(abc\)"), std::string(R"(abc\)"));
  EXPECT_EQ(Tag::UnescapeTag(R"(\)"), std::string(R"(\)"));
  EXPECT_EQ(Tag::UnescapeTag(R"(a\|b\\c)"), std::string(R"(a|b\c)"));
  EXPECT_EQ(Tag::UnescapeTag(R"(\\\\)"), std::string(R"(\\)"));
  EXPECT_EQ(Tag::UnescapeTag(R"(test\value)"), std::string("testvalue"));
  // edge cases :384-435
  EXPECT_EQ(Unescaped("a||b", '|'), std::set<std::string>({"a", "b"}));
  EXPECT_EQ(Unescaped("a|   |b", '|'), std::set<std::string>({"a", "b"}));
  EXPECT_EQ(Tag::ParseSearchTags("b*", '|').status().code(), vks::StatusCode::kInvalidArgument);
  EXPECT_EQ(Tag::ParseSearchTags("*", '|').status().code(), vks::StatusCode::kInvalidArgument);
  EXPECT_EQ(*Tag::ParseSearchTags(R"(tag\)", '|'), std::set<std::string>({R"(tag\)"}));
  EXPECT_EQ(*Tag::ParseSearchTags(R"(\)", '|'), std::set<std::string>({R"(\)"}));
  EXPECT_EQ(Unescaped("日本語|中文", '|'), std::set<std::string>({"日本語", "中文"}));
  EXPECT_TRUE(Tag::ParseSearchTags("", '|')->empty());
  EXPECT_TRUE(Tag::ParseSearchTags("   ", '|')->empty());
  {  // a case-sensitive index keeps spellings apart (tag.cc:81-90)
    Tag index(',', true);
    EXPECT_OK(index.AddRecord("k1", "Red"));
    EXPECT_OK(index.AddRecord("k2", "red"));
    EXPECT_EQ(AsSet(index.Search(QueryTags(&index, "red"), false)), Keys({"k2"}));
    EXPECT_TRUE(QueryTags(&index, "Red").Evaluate("k1"));
    EXPECT_FALSE(QueryTags(&index, "Red").Evaluate("k2"));
  }
}

// ------------------------------------------------------------------------------------------ numeric_index_test.cc
static void NumericIndexCases() {
  {  // SimpleAddModifyRemove :43-83
    Numeric index;
    EXPECT_EQ(*index.AddRecord("key1", "1.5"), RecordResult::kAdded);
    EXPECT_EQ(*index.AddRecord("key2", "2.0"), RecordResult::kAdded);
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 1.0, true, 2.0, true), false)), Keys({"key1", "key2"}));
    EXPECT_EQ(index.AddRecord("key2", "2.0").status().code(), vks::StatusCode::kAlreadyExists);
    EXPECT_EQ(*index.ModifyRecord("key2", "2.1"), RecordResult::kAdded);
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 2.05, true, 2.2, true), false)), Keys({"key2"}));
    EXPECT_EQ(index.ModifyRecord("key5", "2.1").status().code(), vks::StatusCode::kNotFound);
    EXPECT_FALSE(index.IsTracked("key3"));
    EXPECT_OK(index.AddRecord("key3", "3.0"));
    EXPECT_TRUE(index.IsTracked("key3"));
    EXPECT_TRUE(*index.RemoveRecord("key3"));
    EXPECT_TRUE(index.Search(NumericPredicate(&index, 2.5, true, 3.5, true), false).empty());
    EXPECT_FALSE(index.IsTracked("key3"));
    auto res = index.RemoveRecord("key3");
    EXPECT_TRUE(res.ok() && !*res);
    EXPECT_EQ(*index.AddRecord("key5", "aaa"), RecordResult::kInvalidData);
    EXPECT_EQ(*index.ModifyRecord("key5", "aaa"), RecordResult::kInvalidData);
  }
  {  // DetectsInvalidData :87-108
    Numeric index;
    EXPECT_EQ(*index.AddRecord("key1", "not_a_number"), RecordResult::kInvalidData);
    EXPECT_FALSE(index.IsTracked("key1"));
    EXPECT_EQ(*index.AddRecord("key2", "nan"), RecordResult::kInvalidData);
    EXPECT_EQ(*index.AddRecord("key3", ""), RecordResult::kInvalidData);
    EXPECT_EQ(*index.AddRecord("key4", "42"), RecordResult::kAdded);
    EXPECT_TRUE(index.IsTracked("key4"));
    EXPECT_EQ(*index.ModifyRecord("key4", "still_not_a_number"), RecordResult::kInvalidData);
    EXPECT_FALSE(index.IsTracked("key4"));
  }
  {  // RangeSearchInclusiveExclusive :147-191
    Numeric index;
    const char *v[] = {"1.0", "2.0", "2.2", "3.2", "2.0", "2.1"};
    for (int i = 0; i < 6; i++) EXPECT_OK(index.AddRecord("key" + std::to_string(i + 1), v[i]));
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 1.0, true, 2.1, true), false)), Keys({"key1", "key2", "key5", "key6"}));
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 1.0, false, 2.1, true), false)), Keys({"key2", "key5", "key6"}));
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 1.0, false, 2.1, false), false)), Keys({"key2", "key5"}));
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 1.0, false, 3.5, false), false)),
              Keys({"key2", "key3", "key4", "key5", "key6"}));
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 0.0, false, 2.1, false), false)), Keys({"key1", "key2", "key5"}));
    // negate: everything outside the range plus the keys without the field
    EXPECT_EQ(*index.AddRecord("key7", "abc"), RecordResult::kInvalidData);
    EXPECT_EQ(AsSet(index.Search(NumericPredicate(&index, 1.0, false, 2.1, false), true)),
              Keys({"key1", "key3", "key4", "key6", "key7"}));
  }
  EXPECT_TRUE(Numeric::ParseNumber(" 1e3 ").value_or(0) == 1000.0);
  EXPECT_TRUE(Numeric::ParseNumber("-inf").value_or(0) < -1e300);
  EXPECT_FALSE(Numeric::ParseNumber("0x10").has_value());
  EXPECT_FALSE(Numeric::ParseNumber("1.5abc").has_value());
}

static void PredicateCases() {  // predicate.cc:36-39, 332-341, 362-393, 429-520
  Tag tags(',', false);
  Numeric price;
  EXPECT_OK(tags.AddRecord("a", "red,Dark Blue"));
  EXPECT_OK(tags.AddRecord("b", "green"));
  EXPECT_OK(tags.AddRecord("c", "RED"));
  EXPECT_OK(price.AddRecord("a", "10"));
  EXPECT_OK(price.AddRecord("b", "20"));
  EXPECT_OK(price.AddRecord("d", "30"));
  auto red = [&] { return std::make_unique<TagPredicate>(&tags, *Tag::ParseSearchTags("red", '|')); };
  auto cheap = [&] { return std::make_unique<NumericPredicate>(&price, 0.0, true, 15.0, false); };
  EXPECT_TRUE(red()->Evaluate("a") && red()->Evaluate("c") && !red()->Evaluate("b") && !red()->Evaluate("zzz"));
  EXPECT_TRUE(TagPredicate(&tags, *Tag::ParseSearchTags("dark*", '|')).Evaluate("a"));
  EXPECT_TRUE(cheap()->Evaluate("a") && !cheap()->Evaluate("b") && !cheap()->Evaluate("c"));
  ComposedPredicate both(PredicateType::kComposedAnd);
  both.AddChild(red());
  both.AddChild(cheap());
  EXPECT_TRUE(both.Evaluate("a") && !both.Evaluate("c") && !both.Evaluate("b"));
  ComposedPredicate either(PredicateType::kComposedOr);
  either.AddChild(red());
  either.AddChild(cheap());
  EXPECT_TRUE(either.Evaluate("a") && either.Evaluate("c") && !either.Evaluate("b") && !either.Evaluate("d"));
  NegatePredicate not_red(red());
  EXPECT_TRUE(!not_red.Evaluate("a") && not_red.Evaluate("b") && not_red.Evaluate("d") && not_red.Evaluate("nokey"));
  // the end point matches even when the interval is empty otherwise (the formula's second clause)
  EXPECT_TRUE(NumericPredicate(&price, 50.0, true, 20.0, true).Evaluate("b"));
}

// ------------------------------------------------------------------------------------------ device case (B200)
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint64_t Next() {
  g_rng ^= g_rng << 13;
  g_rng ^= g_rng >> 7;
  g_rng ^= g_rng << 17;
  return g_rng;
}
static float Unit() { return (float)((Next() >> 11) * (1.0 / 9007199254740992.0)); }

struct Doc {
  std::vector<float> vec;
  std::string tags;   // "" = no TAG field
  std::string price;  // "" = no NUMERIC field
};

static std::string_view Bytes(const std::vector<float> &v) {
  return std::string_view(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(float));
}

template <typename Index>
static void DeviceBridgeOn(Index *vectors, int dim, int n, bool hnsw) {
  const char *palette[] = {"red", "green", "blue", "Dark Blue", "darkness", "disagree", "dislike", "RED", "a|b"};
  std::vector<Doc> docs(n);
  for (auto &d : docs) {
    d.vec.resize(dim);
    for (auto &x : d.vec) x = Unit() * 2.0f - 1.0f;
    const int nt = (int)(Next() % 4);  // 0..3 tags
    for (int t = 0; t < nt; t++) d.tags += (t ? " , " : "") + std::string(palette[Next() % 9]);
    if (Next() % 5) d.price = std::to_string((int)(Next() % 1000) / 10.0);
  }
  Tag tags(',', false, vectors);
  Numeric price(vectors);
  DeviceFilterEvaluator evaluator(vectors);
  auto key = [](int i) { return "doc:" + std::to_string(i); };
  // attach the three attributes of a key in every order
  for (int i = 0; i < n; i++) {
    const int order = i % 3;
    if (order == 0) EXPECT_OK(vectors->AddRecord(key(i), Bytes(docs[i].vec)));
    EXPECT_OK(tags.AddRecord(key(i), docs[i].tags));
    if (order == 1) EXPECT_OK(vectors->AddRecord(key(i), Bytes(docs[i].vec)));
    EXPECT_OK(price.AddRecord(key(i), docs[i].price));
    if (order == 2) EXPECT_OK(vectors->AddRecord(key(i), Bytes(docs[i].vec)));
  }
  auto tag_leaf = [&](const char *q) { return std::make_unique<TagPredicate>(&tags, *Tag::ParseSearchTags(q, '|')); };
  auto range = [&](double a, bool ia, double b, bool ib) { return std::make_unique<NumericPredicate>(&price, a, ia, b, ib); };
  auto build_roots = [&]() {
    std::vector<std::unique_ptr<Predicate>> roots;
    roots.push_back(tag_leaf("red"));
    roots.push_back(tag_leaf("dark*"));
    roots.push_back(tag_leaf("green|blue|dis*"));
    roots.push_back(tag_leaf(R"(a\|b)"));
    roots.push_back(tag_leaf("no-such-tag"));
    roots.push_back(range(10.0, true, 20.0, false));
    roots.push_back(range(0.0, false, 99.9, true));
    roots.push_back(range(50.0, true, 50.0, true));
    {
      auto p = std::make_unique<ComposedPredicate>(PredicateType::kComposedAnd);
      p->AddChild(tag_leaf("red"));
      p->AddChild(range(0.0, true, 50.0, true));
      roots.push_back(std::move(p));
    }
    {
      auto p = std::make_unique<ComposedPredicate>(PredicateType::kComposedOr);
      p->AddChild(tag_leaf("green"));
      p->AddChild(range(90.0, true, 100.0, true));
      roots.push_back(std::move(p));
    }
    roots.push_back(std::make_unique<NegatePredicate>(tag_leaf("red|green")));
    {
      auto p = std::make_unique<ComposedPredicate>(PredicateType::kComposedAnd);
      p->AddChild(std::make_unique<NegatePredicate>(tag_leaf("blue")));
      p->AddChild(range(25.0, false, 75.0, false));
      p->AddChild(tag_leaf("dis*|dark*"));
      roots.push_back(std::move(p));
    }
    {
      auto inner = std::make_unique<ComposedPredicate>(PredicateType::kComposedAnd);
      inner->AddChild(tag_leaf("red"));
      inner->AddChild(std::make_unique<NegatePredicate>(range(0.0, true, 30.0, true)));
      auto p = std::make_unique<ComposedPredicate>(PredicateType::kComposedOr);
      p->AddChild(std::move(inner));
      p->AddChild(std::make_unique<NegatePredicate>(std::make_unique<NegatePredicate>(tag_leaf("darkness"))));
      roots.push_back(std::move(p));
    }
    return roots;
  };
  std::vector<float> q(dim);
  auto check_all = [&](const char *phase) {
    auto roots = build_roots();
    for (size_t r = 0; r < roots.size(); r++) {
      const Predicate &root = *roots[r];
      std::vector<std::string> want = evaluator.EvaluateOnHost(root);
      auto set = evaluator.Evaluate(root);
      EXPECT_OK(set);
      if (!set.ok()) {
        std::fprintf(stderr, "  %s predicate %zu: %s\n", phase, r, set.status().message().c_str());
        continue;
      }
      const uint64_t bits = (uint64_t)n * 4 + 64;  // every label ever handed out is below this
      std::vector<uint8_t> bitmap((bits + 7) / 8);
      EXPECT_EQ(vkgpu_set_read(vectors->handle(), set->id(), bitmap.data(), bits), 0);
      std::set<std::string> got;
      bool unknown_label = false;
      for (uint64_t l = 0; l < bits; l++)
        if ((bitmap[l >> 3] >> (l & 7)) & 1) {
          auto k = vectors->GetKeyDuringSearch(l);
          if (k.ok()) got.insert(*k); else unknown_label = true;
        }
      EXPECT_FALSE(unknown_label);  // no bit for a label that left the vector index
      EXPECT_EQ(got, std::set<std::string>(want.begin(), want.end()));
      uint64_t card = 0;
      EXPECT_EQ(vkgpu_set_cardinality(vectors->handle(), set->id(), &card), 0);
      EXPECT_EQ(card, (uint64_t)want.size());
      if (got != std::set<std::string>(want.begin(), want.end()))
        std::fprintf(stderr, "  %s predicate %zu: device %zu keys, host %zu keys\n", phase, r, got.size(), want.size());
      // filtered kNN through the device set == the key-list pre-filter over the host-evaluated keys
      for (auto &x : q) x = Unit() * 2.0f - 1.0f;
      auto a = evaluator.Search(Bytes(q), 10, root, hnsw ? std::optional<size_t>(400) : std::nullopt);
      EXPECT_OK(a);
      if (want.empty() && a.ok()) EXPECT_TRUE(a->empty());
      StatusOr<std::vector<Neighbor>> b = std::vector<Neighbor>();
      if constexpr (std::is_same_v<Index, VectorHNSW<float>>) {
        const KeyFilter f = [&](const std::string &k) { return root.Evaluate(k); };
        b = vectors->Search(Bytes(q), 10, CancelNever(), &f, 400);
      } else {
        b = vectors->SearchPrefiltered(Bytes(q), 10, want);
      }
      EXPECT_OK(b);
      if (!a.ok() || !b.ok()) continue;
      EXPECT_EQ(a->size(), b->size());
      if (!hnsw) EXPECT_EQ(a->size(), std::min<size_t>(10, want.size()));
      for (size_t j = 0; j < a->size() && j < b->size(); j++) {
        EXPECT_EQ((*a)[j].external_id, (*b)[j].external_id);
        EXPECT_TRUE(std::memcmp(&(*a)[j].distance, &(*b)[j].distance, 4) == 0);
      }
    }
  };
  check_all("after ingest");
  if constexpr (std::is_same_v<Index, VectorHNSW<float>>) {
    // the planner's other branch (UsePreFiltering, planner.cc:21-46): with the threshold at 1.0 every filtered query on
    // the graph index is answered by exact distances over the qualifying keys — the k nearest of them by brute
    // force in double precision on the host (COSINE index: rank by 1 - cos)
    EXPECT_TRUE(valkey_search::query::UsePreFiltering(1, vectors));
    EXPECT_FALSE(valkey_search::query::UsePreFiltering((size_t)n, vectors));
    evaluator.SetPrefilteringThresholdRatio(1.0);
    auto roots = build_roots();
    for (size_t r = 0; r < roots.size(); r++) {
      std::vector<std::string> want = evaluator.EvaluateOnHost(*roots[r]);
      for (auto &x : q) x = Unit() * 2.0f - 1.0f;
      auto got = evaluator.Search(Bytes(q), 10, *roots[r]);
      EXPECT_OK(got);
      if (!got.ok()) continue;
      std::vector<std::pair<double, std::string>> ranked;
      for (const auto &k : want) {
        const int i = std::atoi(k.c_str() + 4);  // "doc:<i>"
        double dot = 0, nq = 0, nv = 0;
        for (int j = 0; j < dim; j++) {
          dot += (double)q[j] * docs[i].vec[j];
          nq += (double)q[j] * q[j];
          nv += (double)docs[i].vec[j] * docs[i].vec[j];
        }
        ranked.emplace_back(1.0 - dot / std::sqrt(nq * nv), k);
      }
      std::sort(ranked.begin(), ranked.end());
      EXPECT_EQ(got->size(), std::min<size_t>(10, ranked.size()));
      std::set<std::string> a, b;
      for (const auto &nb : *got) a.insert(nb.external_id);
      for (size_t j = 0; j < ranked.size() && j < 10; j++) b.insert(ranked[j].second);
      EXPECT_EQ(a, b);
      for (size_t j = 1; j < got->size(); j++) EXPECT_TRUE((*got)[j - 1].distance <= (*got)[j].distance);
    }
    evaluator.SetPrefilteringThresholdRatio(0.0);  // the checks below compare against the inline filter
  }
  // mutations: vectors leave and come back under new labels, tags and prices change, fields disappear
  for (int i = 0; i < n; i += 7) EXPECT_OK(vectors->RemoveRecord(key(i)));
  for (int i = 0; i < n; i += 14) EXPECT_OK(vectors->AddRecord(key(i), Bytes(docs[i].vec)));
  for (int i = 3; i < n; i += 11) (void)tags.ModifyRecord(key(i), "red, disagree");  // NotFound when the key had no tags
  for (int i = 5; i < n; i += 13) (void)tags.RemoveRecord(key(i));
  for (int i = 1; i < n; i += 9) (void)price.ModifyRecord(key(i), "15.5");
  for (int i = 2; i < n; i += 17) (void)price.RemoveRecord(key(i));
  for (int i = 4; i < n; i += 19) (void)price.ModifyRecord(key(i), "oops");  // invalid: the field is dropped
  check_all("after mutations");
}

// SearchTest (testing/search_test.cc:751-895; the corpus of CreateIndexSchemaWithMultipleAttributes :466-540): 10 000
// vectors v[i][j] = 10 (i + j) / 10100 in 100 dimensions (lower index = closer to the zero query), numeric = i, tags
// "LT10000" (+ ",LT5" for i < 5, + ",LT3" for i < 3); zero query, k = 5, ef = 30; the reference's 15 filters and
// the key SETS it expects, for both index types.  The filters are built as predicate trees (the filter parser is
// command surface and stays in the module).
static const char *g_graph_path = nullptr;  // --graph FILE: an HNSW stream written by the reference for this corpus

template <typename Index>
static void ReferenceSearchTestOn(Index *vectors, bool vectors_loaded = false) {
  constexpr int kDim = 100, kRecords = 10000;
  Tag tag(',', false, vectors);
  Numeric numeric(vectors);
  DeviceFilterEvaluator evaluator(vectors);
  std::vector<float> v(kDim);
  for (int i = 0; i < kRecords; i++) {
    for (int j = 0; j < kDim; j++) v[j] = 10.0f * ((float)(i + j) / (float)(kRecords + kDim));  // testing/common.cc:42-53
    const std::string key = std::to_string(i);
    if (!vectors_loaded) EXPECT_OK(vectors->AddRecord(key, Bytes(v)));
    EXPECT_OK(numeric.AddRecord(key, std::to_string(i)));
    std::string tags = "LT10000";
    if (i < 5) tags += ",LT5";
    if (i < 3) tags += ",LT3";
    EXPECT_OK(tag.AddRecord(key, tags));
  }
  auto T = [&](const char *q) { return std::make_unique<TagPredicate>(&tag, *Tag::ParseSearchTags(q, '|')); };
  auto N = [&](double a, double b) { return std::make_unique<NumericPredicate>(&numeric, a, true, b, true); };
  auto Not = [](std::unique_ptr<Predicate> p) { return std::make_unique<NegatePredicate>(std::move(p)); };
  auto Two = [](PredicateType t, std::unique_ptr<Predicate> a, std::unique_ptr<Predicate> b) {
    auto p = std::make_unique<ComposedPredicate>(t);
    p->AddChild(std::move(a));
    p->AddChild(std::move(b));
    return p;
  };
  struct Case {
    const char *name;
    std::unique_ptr<Predicate> filter;  // null: no filter
    std::set<std::string> expected;
  };
  std::vector<Case> cases;
  const std::set<std::string> first5 = {"0", "1", "2", "3", "4"};
  cases.push_back({"no_filter", nullptr, first5});
  cases.push_back({"prefix_match_filter", T("lT*"), first5});
  cases.push_back({"numeric_filter_all_candidates_eligible", N(0, 10000), first5});
  cases.push_back({"numeric_filter_k_eligible_candidates", N(0, 4), first5});
  cases.push_back({"numeric_filter_less_than_k_eligible_candidates", N(0, 2), {"0", "1", "2"}});
  cases.push_back({"numeric_filter_no_eligible_candidates", N(10000, 20000), {}});
  cases.push_back({"tag_filter_all_candidates_eligible", T("LT10000"), first5});
  cases.push_back({"tag_filter_k_eligible_candidates", T("LT5"), first5});
  cases.push_back({"tag_filter_less_than_k_eligible_candidates", T("LT3"), {"0", "1", "2"}});
  cases.push_back({"tag_filter_no_eligible_candidates", T("random"), {}});
  cases.push_back({"or_filter", Two(PredicateType::kComposedOr, N(4, 100), T("LT5")), first5});
  cases.push_back({"and_filter", Two(PredicateType::kComposedAnd, N(4, 100), T("LT5")), {"4"}});
  cases.push_back({"numeric_negate_filter", Not(N(0, 100)), {"101", "102", "103", "104", "105"}});
  cases.push_back({"tag_negate_filter", Not(T("LT5")), {"5", "6", "7", "8", "9"}});
  cases.push_back({"composite_filter_with_negate", Two(PredicateType::kComposedAnd, Not(N(4, 100)), T("LT5")), {"0", "1", "2", "3"}});
  {
    // FetchFilteredKeysTest (search_test.cc:677-749): CalcBestMatchingPrefilteredKeys with k = 100 over fetched key
    // ranges — every fetched key is re-evaluated against the predicate, a key fetched twice counts once
    struct Fetch {
      std::unique_ptr<Predicate> filter;
      std::vector<std::pair<int, int>> ranges;
      std::set<std::string> expected;
    };
    std::vector<Fetch> fetches;
    fetches.push_back({N(0, 4), {{0, 4}}, {"0", "1", "2", "3", "4"}});
    fetches.push_back({Two(PredicateType::kComposedOr, N(0, 4), N(1, 6)), {{0, 4}, {1, 6}}, {"0", "1", "2", "3", "4", "5", "6"}});
    fetches.push_back({Two(PredicateType::kComposedAnd, N(0, 4), N(1, 6)), {{0, 4}}, {"1", "2", "3", "4"}});
    fetches.push_back({N(1, 5), {{0, 4}}, {"1", "2", "3", "4"}});
    for (int j = 0; j < kDim; j++) v[j] = 10.0f * ((float)j / (float)(1 + kDim));  // DeterministicallyGenerateVectors(1, ...)
    for (const auto &f : fetches) {
      std::vector<std::string> keys;
      for (const auto &range : f.ranges)
        for (int i = range.first; i <= range.second; i++)
          if (f.filter->Evaluate(std::to_string(i))) keys.push_back(std::to_string(i));
      auto r = vectors->SearchPrefiltered(Bytes(v), 100, keys);
      EXPECT_OK(r);
      if (!r.ok()) continue;
      std::set<std::string> got;
      for (const auto &nb : *r) got.insert(nb.external_id);
      EXPECT_EQ(r->size(), f.expected.size());
      EXPECT_EQ(got, f.expected);
    }
  }
  const std::vector<float> zero(kDim, 0.0f);
  for (const auto &c : cases) {
    StatusOr<std::vector<Neighbor>> r = std::vector<Neighbor>();
    if (c.filter) {
      r = evaluator.Search(Bytes(zero), 5, *c.filter, 30);
    } else {
      if constexpr (std::is_same_v<Index, VectorHNSW<float>>)
        r = vectors->Search(Bytes(zero), 5, CancelNever(), nullptr, 30);
      else
        r = vectors->Search(Bytes(zero), 5, CancelNever());
    }
    EXPECT_OK(r);
    if (!r.ok()) continue;
    std::set<std::string> got;
    for (const auto &nb : *r) got.insert(nb.external_id);
    EXPECT_EQ(r->size(), c.expected.size());
    EXPECT_EQ(got, c.expected);
    if (got != c.expected) {
      std::string s;
      for (const auto &k : got) s += k + " ";
      std::fprintf(stderr, "  %s: got { %s}\n", c.name, s.c_str());
    }
  }
}
// LocalSearchTest (testing/search_test.cc:542-676): FLAT index, L2 and COSINE, the same corpus, query = (1, ..., 1):
// how many neighbours each filter yields (vector queries: min(k, qualifying keys); non-vector queries: the qualifying
// keys themselves, here the cardinality of the device-evaluated set) and, for COSINE, distances within [0, 2].
static void ReferenceLocalSearchTest() {
  for (auto metric : {DistanceMetric::kL2, DistanceMetric::kCosine}) {
    constexpr int kDim = 100, kRecords = 10000;
    VectorIndexProto p;
    p.dimension_count = kDim;
    p.distance_metric = metric;
    p.initial_cap = 1000;
    p.flat_algorithm.block_size = 250;
    auto flat = VectorFlat<float>::Create(p);
    EXPECT_OK(flat);
    if (!flat.ok()) return;
    VectorFlat<float> *vectors = flat->get();
    Tag tag(',', false, vectors);
    Numeric numeric(vectors);
    DeviceFilterEvaluator evaluator(vectors);
    std::vector<float> v(kDim);
    for (int i = 0; i < kRecords; i++) {
      for (int j = 0; j < kDim; j++) v[j] = 10.0f * ((float)(i + j) / (float)(kRecords + kDim));
      const std::string key = std::to_string(i);
      EXPECT_OK(vectors->AddRecord(key, Bytes(v)));
      EXPECT_OK(numeric.AddRecord(key, std::to_string(i)));
      EXPECT_OK(tag.AddRecord(key, std::string("LT10000") + (i < 5 ? ",LT5" : "") + (i < 3 ? ",LT3" : "")));
    }
    auto T = [&](const char *q) { return std::make_unique<TagPredicate>(&tag, *Tag::ParseSearchTags(q, '|')); };
    auto N = [&](double a, double b) { return std::make_unique<NumericPredicate>(&numeric, a, true, b, true); };
    auto Two = [](PredicateType t, std::unique_ptr<Predicate> a, std::unique_ptr<Predicate> b) {
      auto c = std::make_unique<ComposedPredicate>(t);
      c->AddChild(std::move(a));
      c->AddChild(std::move(b));
      return c;
    };
    struct Case {
      std::unique_ptr<Predicate> filter;
      int k;  // 0: not a vector query
      size_t expected;
    };
    std::vector<Case> cases;
    cases.push_back({N(10, 20), 10, 10});
    cases.push_back({N(99, 99), 10, 1});
    cases.push_back({N(10000, 20000), 10, 0});
    cases.push_back({T("LT5"), 5, 5});
    cases.push_back({T("LT3"), 5, 3});
    cases.push_back({T("Lt*"), 10, 10});
    cases.push_back({T("random"), 10, 0});
    cases.push_back({N(1, 10), 0, 10});
    cases.push_back({Two(PredicateType::kComposedAnd, N(1, 10), T("LT5")), 0, 4});
    cases.push_back({Two(PredicateType::kComposedOr, N(1, 10), N(21, 25)), 0, 15});
    const std::vector<float> ones(kDim, 1.0f);
    for (const auto &c : cases) {
      if (c.k == 0) {
        auto set = evaluator.Evaluate(*c.filter);
        EXPECT_OK(set);
        if (!set.ok()) continue;
        uint64_t card = 0;
        EXPECT_EQ(vkgpu_set_cardinality(vectors->handle(), set->id(), &card), 0);
        EXPECT_EQ(card, (uint64_t)c.expected);
        EXPECT_EQ(evaluator.EvaluateOnHost(*c.filter).size(), c.expected);
        continue;
      }
      auto r = evaluator.Search(Bytes(ones), c.k, *c.filter);
      EXPECT_OK(r);
      if (!r.ok()) continue;
      EXPECT_EQ(r->size(), c.expected);
      if (metric == DistanceMetric::kCosine)
        for (const auto &nb : *r) EXPECT_TRUE(nb.distance >= 0.0f && nb.distance <= 2.0f);
    }
  }
}

static void ReferenceSearchTestFlat() {
  VectorIndexProto p;
  p.dimension_count = 100;
  p.distance_metric = DistanceMetric::kL2;
  p.initial_cap = 1000;
  p.flat_algorithm.block_size = 250;
  auto flat = VectorFlat<float>::Create(p);
  EXPECT_OK(flat);
  if (flat.ok()) ReferenceSearchTestOn(flat->get());
}
// One chunk stream in a file: u64 chunk count, then per chunk u64 length + bytes (the container of oracle/ref_capi.cc)
struct FileStream : public InputStream {
  std::vector<std::string> chunks;
  size_t next = 0;
  bool Read(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    uint64_t n = 0;
    bool ok = std::fread(&n, 8, 1, f) == 1;
    for (uint64_t i = 0; ok && i < n; i++) {
      uint64_t len = 0;
      ok = std::fread(&len, 8, 1, f) == 1 && len < (1ull << 32);
      if (!ok) break;
      std::string c(len, '\0');
      ok = len == 0 || std::fread(c.data(), 1, len, f) == len;
      chunks.push_back(std::move(c));
    }
    std::fclose(f);
    return ok;
  }
  vks::StatusOr<std::unique_ptr<std::string>> LoadChunk() override {
    if (next >= chunks.size()) return vks::NotFoundError("no more chunks");
    return std::make_unique<std::string>(chunks[next++]);
  }
  bool HasNext() const override { return next < chunks.size(); }
};
static void ReferenceSearchTestHnsw() {
  VectorIndexProto p;
  p.dimension_count = 100;
  p.distance_metric = DistanceMetric::kL2;
  p.initial_cap = 1000;
  p.hnsw_algorithm.m = 10;
  p.hnsw_algorithm.ef_construction = 300;
  p.hnsw_algorithm.ef_runtime = 30;
  if (g_graph_path) {
    // the reference's OWN graph for this corpus (built and saved by its hnswlib through oracle/_ref), loaded the way
    // an RDB load would: the searches below then run on exactly the graph the reference's SearchTest runs on
    FileStream data;
    EXPECT_TRUE(data.Read(g_graph_path));
    auto hnsw = VectorHNSW<float>::LoadFromStream(p, data);
    EXPECT_OK(hnsw);
    if (!hnsw.ok()) return;
    FileStream keys;  // TrackedKeyMetadata: key = decimal label, as the corpus assigns them
    for (int i = 0; i < 10000; i++) {
      TrackedKeyMetadataPb pb;
      pb.key = std::to_string(i);
      pb.internal_id = (uint64_t)i;
      pb.magnitude = kDefaultMagnitude;
      keys.chunks.push_back(pb.SerializeAsString());
    }
    EXPECT_OK((*hnsw)->LoadTrackedKeys(keys));
    ReferenceSearchTestOn(hnsw->get(), true);
    return;
  }
  auto hnsw = VectorHNSW<float>::Create(p);
  EXPECT_OK(hnsw);
  if (hnsw.ok()) ReferenceSearchTestOn(hnsw->get());
}

static void DeviceBridgeFlat() {
  {
    VectorIndexProto p;
    p.dimension_count = 32;
    p.distance_metric = DistanceMetric::kL2;
    p.initial_cap = 4096;
    p.flat_algorithm.block_size = 1024;
    auto flat = VectorFlat<float>::Create(p);
    EXPECT_OK(flat);
    if (flat.ok()) DeviceBridgeOn(flat->get(), 32, 3000, false);
  }
}
static void DeviceBridgeHnsw() {
  {
    VectorIndexProto p;
    p.dimension_count = 32;
    p.distance_metric = DistanceMetric::kCosine;
    p.initial_cap = 2048;
    p.hnsw_algorithm.m = 16;
    p.hnsw_algorithm.ef_construction = 100;
    p.hnsw_algorithm.ef_runtime = 50;
    auto hnsw = VectorHNSW<float>::Create(p);
    EXPECT_OK(hnsw);
    if (hnsw.ok()) DeviceBridgeOn(hnsw->get(), 32, 1200, true);
  }
}

// A hybrid query over a SHARDED index, the way the module would run it on a multi-GPU box:
// every shard is a complete stack — vector index + TAG + NUMERIC + DeviceFilterEvaluator over ITS keys — the root
// predicate is evaluated on each shard into a device set of that shard, and ONE vkgpu_sharded_search_batch over the
// adopted shard handles ranks the qualifying rows of all shards (src/query/search.cc:401-481 on every node +
// src/query/fanout.cc:159-220).  The answer must be the single-stack answer: same keys, same distance bits.
static void DeviceBridgeSharded() {
  const int kDim = 24, kN = 2400, kShards = 3, kK = 12;
  int ndev = vkgpu_device_count();
  if (ndev < 1) ndev = 1;
  struct Stack {
    std::shared_ptr<VectorFlat<float>> vectors;
    std::unique_ptr<Tag> tags;
    std::unique_ptr<Numeric> price;
    std::unique_ptr<DeviceFilterEvaluator> evaluator;
  };
  auto make = [&](int device, uint64_t base) {
    VectorIndexProto p;
    p.dimension_count = kDim;
    p.distance_metric = DistanceMetric::kL2;
    p.initial_cap = kN;
    p.flat_algorithm.block_size = 512;
    p.gpu_device = device;
    p.gpu_label_base = base;
    Stack s;
    auto flat = VectorFlat<float>::Create(p);
    EXPECT_OK(flat);
    if (!flat.ok()) return s;
    s.vectors = *flat;
    s.tags = std::make_unique<Tag>(',', false, s.vectors.get());
    s.price = std::make_unique<Numeric>(s.vectors.get());
    s.evaluator = std::make_unique<DeviceFilterEvaluator>(s.vectors.get());
    return s;
  };
  Stack one = make(0, 0);
  std::vector<Stack> shards;
  for (int g = 0; g < kShards; g++) shards.push_back(make(g % ndev, (uint64_t)g * 100000));
  if (!one.vectors) return;
  const char *palette[] = {"red", "green", "blue", "dark blue"};
  auto key = [](int i) { return "doc:" + std::to_string(i); };
  std::vector<std::vector<float>> vecs(kN, std::vector<float>(kDim));
  for (int i = 0; i < kN; i++) {
    for (auto &x : vecs[i]) x = Unit() * 2.0f - 1.0f;
    if (i % 97 == 5) vecs[i] = vecs[5];  // equal distances across shards: (distance, id) order, ids name the shard
    const std::string tags = std::string(palette[Next() % 4]) + (i % 3 == 0 ? ",sale" : "");
    const std::string price = std::to_string((int)(Next() % 1000) / 10.0);
    for (Stack *s : {&one, &shards[(size_t)(i % kShards)]}) {
      if (!s->vectors) return;
      EXPECT_OK(s->vectors->AddRecord(key(i), Bytes(vecs[i])));
      EXPECT_OK(s->tags->AddRecord(key(i), tags));
      EXPECT_OK(s->price->AddRecord(key(i), price));
    }
  }
  std::vector<vkgpu_index *> handles;
  for (auto &s : shards) handles.push_back(s.vectors->handle());
  vkgpu_sharded *sharded = nullptr;
  EXPECT_EQ(vkgpu_sharded_adopt(handles.data(), (uint32_t)handles.size(), &sharded), 0);
  if (!sharded) return;
  // the same predicate tree built over each stack's own attribute indexes
  auto build = [&](Stack &s, int which) -> std::unique_ptr<Predicate> {
    auto tag_leaf = [&](const char *q) { return std::make_unique<TagPredicate>(s.tags.get(), *Tag::ParseSearchTags(q, '|')); };
    auto range = [&](double a, double b) { return std::make_unique<NumericPredicate>(s.price.get(), a, true, b, false); };
    if (which == 0) return tag_leaf("red");
    if (which == 1) return range(10.0, 30.0);
    if (which == 2) {
      auto p = std::make_unique<ComposedPredicate>(PredicateType::kComposedAnd);
      p->AddChild(tag_leaf("sale"));
      p->AddChild(std::make_unique<NegatePredicate>(tag_leaf("dark*")));
      return p;
    }
    if (which == 3) {
      auto p = std::make_unique<ComposedPredicate>(PredicateType::kComposedOr);
      p->AddChild(tag_leaf("green"));
      p->AddChild(range(95.0, 100.0));
      return p;
    }
    return tag_leaf("no-such-tag");
  };
  const int kQueries = 6;
  std::vector<float> Q((size_t)kQueries * kDim);
  for (auto &x : Q) x = Unit() * 2.0f - 1.0f;
  std::memcpy(Q.data(), vecs[5].data(), (size_t)kDim * 4);  // query 0 sits on the duplicated rows
  for (int which = 0; which < 5; which++) {
    // per shard: predicate -> device set of that shard -> B filters naming it
    std::vector<DeviceSetRef> sets;
    std::vector<std::vector<vkgpu_filter>> filters((size_t)kShards, std::vector<vkgpu_filter>((size_t)kQueries));
    std::vector<const vkgpu_filter *> per_shard;
    bool ok = true;
    for (int g = 0; g < kShards; g++) {
      auto root = build(shards[(size_t)g], which);
      auto set = shards[(size_t)g].evaluator->Evaluate(*root);
      EXPECT_OK(set);
      if (!set.ok()) {
        ok = false;
        break;
      }
      for (auto &f : filters[(size_t)g]) f.device_set = set->id();
      sets.push_back(std::move(*set));
      per_shard.push_back(filters[(size_t)g].data());
    }
    if (!ok) continue;
    std::vector<float> dist((size_t)kQueries * kK);
    std::vector<uint64_t> labels((size_t)kQueries * kK);
    std::vector<uint32_t> n((size_t)kQueries);
    EXPECT_EQ(vkgpu_sharded_search_batch(sharded, Q.data(), kQueries, kK, 0, per_shard.data(), 0, dist.data(), labels.data(),
                                         n.data()), 0);
    auto root1 = build(one, which);
    for (int q = 0; q < kQueries; q++) {
      auto want = one.evaluator->Search(std::string_view(reinterpret_cast<const char *>(&Q[(size_t)q * kDim]), (size_t)kDim * 4),
                                        kK, *root1);
      EXPECT_OK(want);
      if (!want.ok()) continue;
      EXPECT_EQ((size_t)n[(size_t)q], want->size());
      if ((size_t)n[(size_t)q] != want->size()) continue;
      // the merged ids name their shard (disjoint id ranges): translate back to keys through that shard's tracker.
      // Keys at exactly equal distances may come in a different order (ids are assigned per shard): compare as
      // (distance bits, key) multisets, which pins every rank outside a tie and the membership inside one.
      std::multiset<std::pair<uint32_t, std::string>> got, exp;
      for (uint32_t i = 0; i < n[(size_t)q]; i++) {
        const uint64_t id = labels[(size_t)q * kK + i];
        const size_t g = (size_t)(id / 100000);
        EXPECT_TRUE(g < (size_t)kShards);
        if (g >= (size_t)kShards) continue;
        auto k2 = shards[g].vectors->GetKeyDuringSearch(id);
        EXPECT_OK(k2);
        uint32_t bits;
        std::memcpy(&bits, &dist[(size_t)q * kK + i], 4);
        if (k2.ok()) got.emplace(bits, *k2);
        if (i) EXPECT_TRUE(dist[(size_t)q * kK + i] >= dist[(size_t)q * kK + i - 1]);
      }
      for (const auto &nb : *want) {
        uint32_t bits;
        std::memcpy(&bits, &nb.distance, 4);
        exp.emplace(bits, nb.external_id);
      }
      // a tie at the k-th place may legitimately keep different keys: drop the last distance class before comparing
      if (!exp.empty() && exp.size() == (size_t)kK) {
        const uint32_t last = std::prev(exp.end())->first;
        for (auto it = exp.begin(); it != exp.end();) it = it->first == last ? exp.erase(it) : std::next(it);
        for (auto it = got.begin(); it != got.end();) it = it->first == last ? got.erase(it) : std::next(it);
      }
      EXPECT_TRUE(got == exp);
    }
  }
  // rows are added through the shards' own handles, not through an adopted handle
  EXPECT_EQ(vkgpu_sharded_remove(sharded, 1), (int)VKGPU_ERR_UNSUPPORTED);
  vkgpu_sharded_destroy(sharded);
}

// --tag-golden FILE: a TAG index and queries from a text file (every string hex-encoded: "sep XX", "case 0|1",
// "doc KEY TAGS", "query TEXT" with TEXT = the whole filter expression "@field:{ ... }"); prints per query
// "N key key ..." (hex keys, sorted) or "ERR message".  Host only.  tests/test_zz_filter_bridge.py feeds it the
// RediSearch answers recorded by the reference's compatibility suite.
static std::string FromHex(const std::string &h) {
  std::string out;
  for (size_t i = 0; i + 1 < h.size(); i += 2) out.push_back((char)std::stoi(h.substr(i, 2), nullptr, 16));
  return out;
}
static std::string ToHex(const std::string &s) {
  static const char *d = "0123456789abcdef";
  std::string out;
  for (unsigned char c : s) {
    out.push_back(d[c >> 4]);
    out.push_back(d[c & 15]);
  }
  return out;
}
static int TagGoldenMode(const char *path) {
  FILE *f = std::fopen(path, "r");
  if (!f) return 2;
  char sep = ',';
  bool case_sensitive = false;
  std::unique_ptr<Tag> index;
  std::vector<char> line(1 << 16);
  while (std::fgets(line.data(), (int)line.size(), f)) {
    std::string l(line.data());
    while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
    const size_t sp = l.find(' ');
    const std::string cmd = l.substr(0, sp), rest = sp == std::string::npos ? "" : l.substr(sp + 1);
    if (cmd == "sep") {
      sep = FromHex(rest)[0];
    } else if (cmd == "case") {
      case_sensitive = rest == "1";
    } else if (cmd == "doc") {
      if (!index) index = std::make_unique<Tag>(sep, case_sensitive);
      const size_t sp2 = rest.find(' ');
      const std::string key = FromHex(rest.substr(0, sp2)), tags = sp2 == std::string::npos ? "" : FromHex(rest.substr(sp2 + 1));
      if (!index->AddRecord(key, tags).ok()) return 3;
    } else if (cmd == "query") {
      if (!index) index = std::make_unique<Tag>(sep, case_sensitive);
      const std::string q = FromHex(rest);
      const size_t brace = q.find('{');
      if (brace == std::string::npos) {
        std::printf("ERR no tag clause\n");
        continue;
      }
      auto tag_string = Tag::ParseTagString(std::string_view(q).substr(brace + 1));
      if (!tag_string.ok()) {
        std::printf("ERR %s\n", tag_string.status().message().c_str());
        continue;
      }
      auto parsed = Tag::ParseSearchTags(*tag_string, '|');  // FilterParser::ParseQueryTags
      if (!parsed.ok()) {
        std::printf("ERR %s\n", parsed.status().message().c_str());
        continue;
      }
      TagPredicate predicate(index.get(), *parsed);
      // candidates from the index (Tag::Search), each confirmed by the predicate, de-duplicated: the pre-filter loop
      std::set<std::string> keys;
      for (const auto &k : index->Search(predicate, false))
        if (predicate.Evaluate(k)) keys.insert(k);
      std::printf("%zu", keys.size());
      for (const auto &k : keys) std::printf(" %s", ToHex(k).c_str());
      std::printf("\n");
    }
  }
  std::fclose(f);
  return 0;
}

// --eval FILE: indexes, records, a universe of keys and predicate trees from a text file; prints, per tree, the keys
// of the universe the tree accepts (hex, sorted) — Predicate::Evaluate per key, the reference's pre-filter loop.
// tests/test_filter_oracle.py compares the output with oracle/filter_oracle.py on random inputs.  Host only.
//   tagindex NAME SEPHEX 0|1      numindex NAME
//   tadd|tmod NAME KEYHEX DATAHEX   trem NAME KEYHEX none|record     (same with n... for numeric)
//   universe KEYHEX
//   pred TOKENS...   prefix notation: AND n | OR n | NOT | TAG NAME TAGSTRINGHEX | NUM NAME start incl end incl
struct EvalState {
  std::map<std::string, std::unique_ptr<Tag>> tags;
  std::map<std::string, std::unique_ptr<Numeric>> nums;
  std::vector<std::string> universe;
};
static std::unique_ptr<Predicate> BuildPredicate(EvalState &st, std::vector<std::string> &tok, size_t &pos, std::string &err) {
  if (pos >= tok.size()) {
    err = "truncated predicate";
    return nullptr;
  }
  const std::string t = tok[pos++];
  if (t == "AND" || t == "OR") {
    const int n = std::atoi(tok[pos++].c_str());
    auto p = std::make_unique<ComposedPredicate>(t == "AND" ? PredicateType::kComposedAnd : PredicateType::kComposedOr);
    for (int i = 0; i < n; i++) {
      auto c = BuildPredicate(st, tok, pos, err);
      if (!c) return nullptr;
      p->AddChild(std::move(c));
    }
    return p;
  }
  if (t == "NOT") {
    auto c = BuildPredicate(st, tok, pos, err);
    if (!c) return nullptr;
    return std::make_unique<NegatePredicate>(std::move(c));
  }
  if (t == "TAG") {
    Tag *ix = st.tags.at(tok[pos++]).get();
    auto parsed = Tag::ParseSearchTags(FromHex(tok[pos++]), '|');
    if (!parsed.ok()) {
      err = parsed.status().message();
      return nullptr;
    }
    return std::make_unique<TagPredicate>(ix, *parsed);
  }
  if (t == "NUM") {
    Numeric *ix = st.nums.at(tok[pos++]).get();
    const double a = std::strtod(tok[pos++].c_str(), nullptr);
    const bool ia = tok[pos++] == "1";
    const double b = std::strtod(tok[pos++].c_str(), nullptr);
    const bool ib = tok[pos++] == "1";
    return std::make_unique<NumericPredicate>(ix, a, ia, b, ib);
  }
  err = "unknown token " + t;
  return nullptr;
}
static int EvalMode(const char *path) {
  FILE *f = std::fopen(path, "r");
  if (!f) return 2;
  EvalState st;
  std::vector<char> line(1 << 20);
  while (std::fgets(line.data(), (int)line.size(), f)) {
    std::vector<std::string> tok;
    {
      std::string cur;
      for (const char *p = line.data(); *p; p++) {
        if (*p == ' ' || *p == '\n' || *p == '\r') {
          if (!cur.empty()) tok.push_back(cur);
          cur.clear();
        } else {
          cur.push_back(*p);
        }
      }
      if (!cur.empty()) tok.push_back(cur);
    }
    if (tok.empty()) continue;
    const std::string &cmd = tok[0];
    auto arg = [&](size_t i) { return i < tok.size() ? tok[i] : std::string(); };
    auto report = [&](const vks::StatusOr<RecordResult> &r) {
      if (!r.ok()) std::printf("rec ERR %s\n", r.status().message().c_str());
      else std::printf("rec %s\n", *r == RecordResult::kAdded ? "added" : *r == RecordResult::kMissing ? "missing" : "invalid");
    };
    if (cmd == "tagindex") {
      st.tags[arg(1)] = std::make_unique<Tag>(FromHex(arg(2))[0], arg(3) == "1");
    } else if (cmd == "numindex") {
      st.nums[arg(1)] = std::make_unique<Numeric>();
    } else if (cmd == "tadd") {
      report(st.tags.at(arg(1))->AddRecord(FromHex(arg(2)), FromHex(arg(3))));
    } else if (cmd == "tmod") {
      report(st.tags.at(arg(1))->ModifyRecord(FromHex(arg(2)), FromHex(arg(3))));
    } else if (cmd == "trem") {
      auto r = st.tags.at(arg(1))->RemoveRecord(FromHex(arg(2)), arg(3) == "record" ? DeletionType::kRecord : DeletionType::kNone);
      std::printf("rem %d\n", r.ok() && *r ? 1 : 0);
    } else if (cmd == "nadd") {
      report(st.nums.at(arg(1))->AddRecord(FromHex(arg(2)), FromHex(arg(3))));
    } else if (cmd == "nmod") {
      report(st.nums.at(arg(1))->ModifyRecord(FromHex(arg(2)), FromHex(arg(3))));
    } else if (cmd == "nrem") {
      auto r = st.nums.at(arg(1))->RemoveRecord(FromHex(arg(2)), arg(3) == "record" ? DeletionType::kRecord : DeletionType::kNone);
      std::printf("rem %d\n", r.ok() && *r ? 1 : 0);
    } else if (cmd == "universe") {
      st.universe.push_back(FromHex(arg(1)));
    } else if (cmd == "pred") {
      size_t pos = 1;
      std::string err;
      auto p = BuildPredicate(st, tok, pos, err);
      if (!p) {
        std::printf("pred ERR %s\n", err.c_str());
        continue;
      }
      std::set<std::string> keys;
      for (const auto &k : st.universe)
        if (p->Evaluate(k)) keys.insert(k);
      std::printf("pred %zu", keys.size());
      for (const auto &k : keys) std::printf(" %s", ToHex(k).c_str());
      std::printf("\n");
    }
  }
  std::fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 2 && std::string(argv[1]) == "--tag-golden") return TagGoldenMode(argv[2]);
  if (argc > 2 && std::string(argv[1]) == "--eval") return EvalMode(argv[2]);
  const bool host_only = argc > 1 && std::string(argv[1]) == "--host-only";
  struct Case {
    const char *name;
    void (*fn)();
    bool needs_gpu;
  } cases[] = {{"TagIndex", TagIndexCases, false},
               {"NumericIndex", NumericIndexCases, false},
               {"Predicates", PredicateCases, false},
               {"DeviceBridgeFlat", DeviceBridgeFlat, true},
               {"DeviceBridgeHnsw", DeviceBridgeHnsw, true},
               {"DeviceBridgeSharded", DeviceBridgeSharded, true},
               {"ReferenceSearchTestFlat", ReferenceSearchTestFlat, true},
               {"ReferenceLocalSearchTest", ReferenceLocalSearchTest, true},
               {"ReferenceSearchTestHnsw", ReferenceSearchTestHnsw, true}};
  setvbuf(stdout, nullptr, _IOLBF, 0);
  const std::string only = argc > 2 && std::string(argv[1]) == "--case" ? argv[2] : "";
  if (argc > 4 && std::string(argv[3]) == "--graph") g_graph_path = argv[4];
  for (const auto &c : cases) {
    if (c.needs_gpu && host_only) continue;
    if (!only.empty() && only != c.name) continue;
    const int before = g_failures;
    c.fn();
    std::printf("[%s] %s\n", g_failures == before ? "  OK  " : "FAILED", c.name);
  }
  std::printf("%d checks, %d failures\n", g_checks, g_failures);
  return g_failures ? 1 : 0;
}
