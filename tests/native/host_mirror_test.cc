// Parity tests for the C++ host mirror (valkey_search_b200/host/vector_index.h), written to read like the
// reference's own testing/vector_test.cc: same fixtures (DeterministicallyGenerateVectors, testing/common.cc:42-53;
// kDimensions/kInitialCap/kBlockSize/kM/kEFConstruction/kEFRuntime, vector_test.cc:54-59; IndexToKey :126-128),
// same cases (TestIndex :238-291 through BasicHNSW/BasicFlat :353-375, EfRuntimeRecall :439-500) plus the
// integration test's cosine goldens (testing/integration/vector_search_integration_test.py:19-23,144-166) and the
// pre-filter path.  Needs a B200: there is no CPU path behind the ABI.  `--host-only` runs the host-side cases.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../valkey_search_b200/host/fanout_merge.h"
#include "../../valkey_search_b200/host/hnsw_serialization.h"
#include "../../valkey_search_b200/host/vector_index.h"

using namespace valkey_search::indexes;

static int g_failures = 0, g_checks = 0;
#define EXPECT_TRUE(c)                                                              \
  do {                                                                              \
    g_checks++;                                                                     \
    if (!(c)) {                                                                     \
      g_failures++;                                                                 \
      std::fprintf(stderr, "  FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);         \
    }                                                                               \
  } while (0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_OK(s) EXPECT_TRUE((s).ok())

constexpr int kDimensions = 100;
constexpr int kInitialCap = 15000;
constexpr uint32_t kBlockSize = 250;
constexpr int kM = 16;
constexpr int kEFConstruction = 20;
constexpr int kEFRuntime = 20;

static std::vector<std::vector<float>> DeterministicallyGenerateVectors(int size, int dimensions, float max_value) {
  std::vector<std::vector<float>> result(size, std::vector<float>(dimensions));
  for (int i = 0; i < size; ++i)
    for (int j = 0; j < dimensions; ++j)
      result[i][j] = max_value * (static_cast<float>(i + j) / static_cast<float>(size + dimensions));
  return result;
}
static std::string IndexToKey(int i) { return std::to_string(i) + "_key"; }
static std::string_view VectorToStr(const std::vector<float> &v) {
  return std::string_view(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(float));
}
static VectorIndexProto CreateHNSWVectorIndexProto(int dim, DistanceMetric metric, int initial_cap, int m, int efc, int ef) {
  VectorIndexProto p;
  p.dimension_count = dim;
  p.distance_metric = metric;
  p.initial_cap = initial_cap;
  p.hnsw_algorithm.m = m;
  p.hnsw_algorithm.ef_construction = efc;
  p.hnsw_algorithm.ef_runtime = ef;
  return p;
}
static VectorIndexProto CreateFlatVectorIndexProto(int dim, DistanceMetric metric, int initial_cap, uint32_t block_size) {
  VectorIndexProto p;
  p.dimension_count = dim;
  p.distance_metric = metric;
  p.initial_cap = initial_cap;
  p.flat_algorithm.block_size = block_size;
  return p;
}

enum class ExpectedResults { kSuccess, kMissing, kInvalidData, kError };

static void VerifyResult(const vks::StatusOr<RecordResult> &res, ExpectedResults expected) {
  if (expected == ExpectedResults::kSuccess) {
    EXPECT_OK(res);
    if (res.ok()) EXPECT_EQ(res.value(), RecordResult::kAdded);
  } else if (expected == ExpectedResults::kMissing) {
    EXPECT_OK(res);
    if (res.ok()) EXPECT_EQ(res.value(), RecordResult::kMissing);
  } else if (expected == ExpectedResults::kInvalidData) {
    EXPECT_OK(res);
    if (res.ok()) EXPECT_EQ(res.value(), RecordResult::kInvalidData);
  } else {
    EXPECT_FALSE(res.status().ok());
  }
}
static void VerifyAdd(VectorBase *index, const std::vector<std::vector<float>> &vectors, int i, ExpectedResults expected) {
  const auto id = IndexToKey(i);
  const bool already = index->IsTracked(id);
  auto res = index->AddRecord(id, VectorToStr(vectors[i]));
  if (res.ok() && res.value() == RecordResult::kAdded) {
    EXPECT_TRUE(index->IsTracked(id));
  } else if (!already) {
    EXPECT_FALSE(index->IsTracked(id));
  }
  VerifyResult(res, expected);
}
static void VerifyModify(VectorBase *index, const std::vector<float> &vector, int i, ExpectedResults expected,
                         bool expected_tracked) {
  const auto id = IndexToKey(i);
  auto res = index->ModifyRecord(id, VectorToStr(vector));
  EXPECT_EQ(index->IsTracked(id), expected_tracked);
  VerifyResult(res, expected);
}

// testing/vector_test.cc:238-291
template <typename T>
static void TestIndex(T *index, int dimensions, int vector_size) {
  auto vectors = DeterministicallyGenerateVectors(vector_size, dimensions, 10.0);
  for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index, vectors, i, ExpectedResults::kSuccess);
  VerifyAdd(index, vectors, 0, ExpectedResults::kError);  // "Embedding id already exists"
  auto vectors_small_dim = DeterministicallyGenerateVectors(vectors.size(), dimensions - 1, 1.0);
  VerifyAdd(index, vectors_small_dim, 0, ExpectedResults::kInvalidData);
  VerifyModify(index, vectors_small_dim[0], 0, ExpectedResults::kInvalidData, false);  // and the key is removed
  VerifyModify(index, vectors[0], 0, ExpectedResults::kError, false);
  VerifyModify(index, vectors[0], vectors.size(), ExpectedResults::kError, false);
  VerifyModify(index, vectors[vectors.size() - 2], vectors.size() - 1, ExpectedResults::kSuccess, true);
  // (ours) an unchanged vector is a no-op reported as kMissing (vector_base.cc:239-243)
  VerifyModify(index, vectors[vectors.size() - 2], vectors.size() - 1, ExpectedResults::kMissing, true);

  for (size_t i = 1; i < vectors.size() - 1; ++i) {
    auto res = index->Search(VectorToStr(vectors[i]), 10, CancelNever());
    EXPECT_OK(res);
    if (res.ok()) {
      EXPECT_FALSE(res->empty());
      bool found = false;
      for (const auto &neighbor : res.value()) {
        if (neighbor.external_id == IndexToKey(i)) {
          EXPECT_TRUE(neighbor.distance - res.value()[0].distance < 0.0001);
          found = true;
          break;
        }
      }
      EXPECT_TRUE(found);
    }
  }
  EXPECT_OK(index->RemoveRecord(IndexToKey(vectors.size()), DeletionType::kNone));
  EXPECT_FALSE(index->RemoveRecord(IndexToKey(vectors.size()), DeletionType::kNone).value());
  for (size_t i = 0; i < vectors.size(); ++i) {
    EXPECT_OK(index->RemoveRecord(IndexToKey(i), DeletionType::kNone));
    EXPECT_FALSE(index->IsTracked(IndexToKey(i)));
  }
  for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index, vectors, i, ExpectedResults::kSuccess);
}

static void BasicHNSW() {
  for (auto metric : {DistanceMetric::kCosine, DistanceMetric::kL2}) {
    auto index = VectorHNSW<float>::Create(
        CreateHNSWVectorIndexProto(kDimensions, metric, kInitialCap, kM, kEFConstruction, kEFRuntime));
    EXPECT_OK(index);
    if (index.ok()) TestIndex<VectorHNSW<float>>(index->get(), kDimensions, 100);
  }
}
static void BasicFlat() {
  for (auto metric : {DistanceMetric::kCosine, DistanceMetric::kL2}) {
    auto index = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, metric, kInitialCap, kBlockSize));
    EXPECT_OK(index);
    if (index.ok()) TestIndex<VectorFlat<float>>(index->get(), kDimensions, 100);
  }
}

// testing/vector_test.cc:418-437
static float CalcRecall(VectorFlat<float> *flat_index, VectorHNSW<float> *hnsw_index, uint64_t k, int dimensions,
                        std::optional<size_t> ef_runtime) {
  auto search_vectors = DeterministicallyGenerateVectors(50, dimensions, 1.5);
  int cnt = 0;
  for (const auto &search_vector : search_vectors) {
    auto res_hnsw = hnsw_index->Search(VectorToStr(search_vector), k, CancelNever(), nullptr, ef_runtime);
    auto res_flat = flat_index->Search(VectorToStr(search_vector), k, CancelNever());
    EXPECT_OK(res_hnsw);
    EXPECT_OK(res_flat);
    if (!res_hnsw.ok() || !res_flat.ok()) return 0.f;
    for (auto &label : *res_hnsw)
      for (auto &real_label : *res_flat)
        if (label.external_id == real_label.external_id) {
          ++cnt;
          break;
        }
  }
  return ((float)cnt) / ((float)(k * search_vectors.size()));
}
// testing/vector_test.cc:439-500
static void EfRuntimeRecall() {
  const int initial_cap = 31000;
  auto index_hnsw = VectorHNSW<float>::Create(
      CreateHNSWVectorIndexProto(kDimensions, DistanceMetric::kL2, initial_cap, kM, kEFConstruction, kEFRuntime));
  auto index_flat = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, DistanceMetric::kL2, initial_cap, kBlockSize));
  EXPECT_OK(index_hnsw);
  EXPECT_OK(index_flat);
  if (!index_hnsw.ok() || !index_flat.ok()) return;
  auto vectors = DeterministicallyGenerateVectors(1000, kDimensions, 2.2);
  for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index_hnsw->get(), vectors, i, ExpectedResults::kSuccess);
  for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index_flat->get(), vectors, i, ExpectedResults::kSuccess);
  const uint64_t k = 10;
  auto no_ef = CalcRecall(index_flat->get(), index_hnsw->get(), k, kDimensions, std::nullopt);
  auto default_ef = CalcRecall(index_flat->get(), index_hnsw->get(), k, kDimensions, kEFRuntime);
  auto ef8 = CalcRecall(index_flat->get(), index_hnsw->get(), k, kDimensions, kEFRuntime * 8);
  std::fprintf(stderr, "  recall@10: ef=default %.3f, ef=%d %.3f, ef=%d %.3f\n", no_ef, kEFRuntime, default_ef,
               kEFRuntime * 8, ef8);
  EXPECT_TRUE(ef8 >= 0.96f);
  EXPECT_EQ(default_ef, no_ef);
}

// testing/integration/vector_search_integration_test.py:19-23,144-166: vectors [1, data, 0...] D=100 COSINE,
// query [1,0,...], k=3 => keys 0,1,2 with scores "0", "0.292893230915", "0.552786409855" (%.12g of the float,
// src/commands/ft_search.cc:69).
static void IntegrationCosineGoldens() {
  for (int algo = 0; algo < 2; algo++) {
    std::shared_ptr<VectorBase> index;
    if (algo == 0) {
      auto r = VectorFlat<float>::Create(CreateFlatVectorIndexProto(100, DistanceMetric::kCosine, 1000, 1024));
      EXPECT_OK(r);
      if (!r.ok()) continue;
      index = *r;
    } else {
      auto r = VectorHNSW<float>::Create(CreateHNSWVectorIndexProto(100, DistanceMetric::kCosine, 1000, 16, 200, 10));
      EXPECT_OK(r);
      if (!r.ok()) continue;
      index = *r;
    }
    for (int i = 0; i < 10; i++) {
      std::vector<float> v(100, 0.0f);
      v[0] = 1.0f;
      v[1] = (float)i;
      EXPECT_OK(index->AddRecord(std::to_string(i), VectorToStr(v)));
    }
    std::vector<float> q(100, 0.0f);
    q[0] = 1.0f;
    vks::StatusOr<std::vector<Neighbor>> res = algo == 0
        ? static_cast<VectorFlat<float> *>(index.get())->Search(VectorToStr(q), 3, CancelNever())
        : static_cast<VectorHNSW<float> *>(index.get())->Search(VectorToStr(q), 3, CancelNever());
    EXPECT_OK(res);
    if (!res.ok()) continue;
    EXPECT_EQ(res->size(), (size_t)3);
    const char *want[3] = {"0", "0.292893230915", "0.552786409855"};
    for (size_t j = 0; j < res->size() && j < 3; j++) {
      char buf[64];
      std::snprintf(buf, sizeof(buf), "%.12g", (*res)[j].distance);
      EXPECT_EQ((*res)[j].external_id, std::to_string(j));
      EXPECT_TRUE(std::strcmp(buf, want[j]) == 0);
      if (std::strcmp(buf, want[j]) != 0) std::fprintf(stderr, "  score[%zu] = %s, want %s\n", j, buf, want[j]);
    }
    // GetValue de-normalises with the stored magnitude (vector_base.cc:279-297)
    auto val = index->GetValue("3");
    EXPECT_OK(val);
    if (val.ok()) {
      const float *f = reinterpret_cast<const float *>(val->data());
      EXPECT_TRUE(std::fabs(f[0] - 1.0f) < 1e-5f && std::fabs(f[1] - 3.0f) < 1e-5f);
    }
  }
}

// AddPrefilteredKey loop (src/query/search.cc:457-481) vs the one-call SearchPrefiltered vs Search with a filter
static void Prefilter() {
  auto r = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, DistanceMetric::kL2, kInitialCap, kBlockSize));
  EXPECT_OK(r);
  if (!r.ok()) return;
  auto index = *r;
  auto vectors = DeterministicallyGenerateVectors(500, kDimensions, 10.0);
  for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index.get(), vectors, i, ExpectedResults::kSuccess);
  std::vector<std::string> keys;
  for (int i = 0; i < 500; i += 7) keys.push_back(IndexToKey(i));
  keys.push_back("not_indexed_key");  // skipped, vector_base.cc:513-516
  auto query = VectorToStr(vectors[250]);
  std::priority_queue<std::pair<float, uint64_t>> results;
  std::unordered_set<std::string> top_keys;
  for (const auto &key : keys)
    if (index->AddPrefilteredKey(query, 5, key, results, top_keys)) top_keys.insert(key);
  auto loop = index->CreateReply(results);
  auto one = index->SearchPrefiltered(query, 5, keys);
  KeyFilter filter = [](const std::string &key) { return std::stoi(key) % 7 == 0; };
  auto via_filter = index->Search(query, 5, CancelNever(), &filter);
  EXPECT_OK(loop);
  EXPECT_OK(one);
  EXPECT_OK(via_filter);
  if (!loop.ok() || !one.ok() || !via_filter.ok()) return;
  EXPECT_EQ(loop->size(), (size_t)5);
  EXPECT_EQ(one->size(), loop->size());
  EXPECT_EQ(via_filter->size(), loop->size());
  for (size_t j = 0; j < loop->size() && j < one->size() && j < via_filter->size(); j++) {
    EXPECT_EQ((*loop)[j].external_id, (*one)[j].external_id);
    EXPECT_TRUE(std::memcmp(&(*loop)[j].distance, &(*one)[j].distance, 4) == 0);
    EXPECT_EQ((*via_filter)[j].external_id, (*one)[j].external_id);
  }
  EXPECT_EQ(top_keys.size(), (size_t)5);
}

// HNSW inline filter (hnswalg.h:515-524) and the batched entry
static void InlineFilterAndBatch() {
  auto rh = VectorHNSW<float>::Create(CreateHNSWVectorIndexProto(kDimensions, DistanceMetric::kL2, kInitialCap, kM, 100, 64));
  auto rf = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, DistanceMetric::kL2, kInitialCap, kBlockSize));
  EXPECT_OK(rh);
  EXPECT_OK(rf);
  if (!rh.ok() || !rf.ok()) return;
  auto vectors = DeterministicallyGenerateVectors(800, kDimensions, 10.0);
  for (size_t i = 0; i < vectors.size(); ++i) {
    VerifyAdd(rh->get(), vectors, i, ExpectedResults::kSuccess);
    VerifyAdd(rf->get(), vectors, i, ExpectedResults::kSuccess);
  }
  KeyFilter even = [](const std::string &key) { return std::stoi(key) % 2 == 0; };
  auto res = (*rh)->Search(VectorToStr(vectors[401]), 10, CancelNever(), &even, 200);
  EXPECT_OK(res);
  if (res.ok()) {
    EXPECT_EQ(res->size(), (size_t)10);
    for (const auto &n : *res) EXPECT_TRUE(std::stoi(n.external_id) % 2 == 0);
  }
  // 64 queries in one launch == 64 single searches
  std::vector<float> flatq;
  for (int b = 0; b < 64; b++) flatq.insert(flatq.end(), vectors[b * 3].begin(), vectors[b * 3].end());
  auto batch = (*rf)->SearchBatch(std::string_view(reinterpret_cast<const char *>(flatq.data()), flatq.size() * 4), 64, 10);
  EXPECT_OK(batch);
  if (batch.ok())
    for (int b = 0; b < 64; b += 9) {
      auto single = (*rf)->Search(VectorToStr(vectors[b * 3]), 10, CancelNever());
      EXPECT_OK(single);
      if (!single.ok()) continue;
      EXPECT_EQ(single->size(), (*batch)[b].size());
      for (size_t j = 0; j < single->size() && j < (*batch)[b].size(); j++) {
        EXPECT_EQ((*single)[j].external_id, (*batch)[b][j].external_id);
        EXPECT_TRUE(std::memcmp(&(*single)[j].distance, &(*batch)[b][j].distance, 4) == 0);
      }
    }
}

// N4, remote-mode requests: in a cluster every shard node answers SearchIndexPartition requests that the coordinators
// of OTHER nodes send it (src/coordinator/server.cc -> query::SearchAsync -> VectorBase::Search, the same local path),
// many at once and each with its own query, and the coordinator that asked folds the replies
// (SearchPartitionResultsTracker, src/query/fanout.cc:153-212).  Here: three shard indexes with the library's dynamic
// batcher on, 48 concurrent request handlers (16 queries x 3 shards, one thread each — the gRPC threads), one tracker
// per query filled in whatever order the shards answer.  Every folded reply must carry the single-index answer's
// distances, and its keys wherever a distance is unique in the reply (the fixture has equal distances, where the fold's
// own tie rule decides), and the batchers must have coalesced the concurrent requests.
static void RemoteModeFanout() {
  using valkey_search::query::fanout::SearchPartitionResultsTracker;
  constexpr int kShards = 3, kQueries = 16;
  const uint64_t k = 10;
  auto vectors = DeterministicallyGenerateVectors(1500, kDimensions, 10.0);
  auto whole = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, DistanceMetric::kL2, 2000, kBlockSize));
  EXPECT_OK(whole);
  if (!whole.ok()) return;
  std::vector<std::shared_ptr<VectorFlat<float>>> shards;
  for (int s = 0; s < kShards; s++) {
    auto proto = CreateFlatVectorIndexProto(kDimensions, DistanceMetric::kL2, 2000, kBlockSize);
    proto.gpu_batch_window_us = 2000;  // the reader pool's batching window
    proto.gpu_max_batch = 64;
    auto r = VectorFlat<float>::Create(proto);
    EXPECT_OK(r);
    if (!r.ok()) return;
    shards.push_back(*r);
  }
  for (size_t i = 0; i < vectors.size(); ++i) {
    VerifyAdd(whole->get(), vectors, i, ExpectedResults::kSuccess);
    VerifyAdd(shards[i % kShards].get(), vectors, i, ExpectedResults::kSuccess);  // keys are cluster-wide unique
  }
  std::vector<SearchPartitionResultsTracker> trackers;
  for (int q = 0; q < kQueries; q++) trackers.emplace_back(k);
  std::vector<std::mutex> mu(kQueries);  // fanout.cc guards its tracker with a mutex: replies arrive on gRPC threads
  std::vector<int> failed(kQueries * kShards, 0);
  std::vector<std::thread> handlers;
  for (int q = 0; q < kQueries; q++)
    for (int s = 0; s < kShards; s++)
      handlers.emplace_back([&, q, s] {
        auto reply = shards[s]->Search(VectorToStr(vectors[q * 7 + 3]), k, CancelNever());  // the shard's local top-k
        if (!reply.ok()) {
          failed[q * kShards + s] = 1;
          return;
        }
        std::lock_guard<std::mutex> lk(mu[q]);
        trackers[q].AddResults(*reply);
      });
  for (auto &t : handlers) t.join();
  for (int f : failed) EXPECT_EQ(f, 0);
  for (int q = 0; q < kQueries; q++) {
    auto want = (*whole)->Search(VectorToStr(vectors[q * 7 + 3]), k, CancelNever());
    EXPECT_OK(want);
    if (!want.ok()) continue;
    auto got = trackers[q].TakeNeighbors();
    EXPECT_EQ(got.size(), want->size());
    for (size_t j = 0; j < got.size() && j < want->size(); j++) {
      EXPECT_TRUE(std::memcmp(&got[j].distance, &(*want)[j].distance, 4) == 0);
      // inside a run of equal distances the coordinator's fold orders by key, descending, and admits nothing equal to
      // its worst once full (fanout_merge.h) — only an entry whose distance is unique in the reply names one key
      const bool tied = (j > 0 && got[j - 1].distance == got[j].distance) ||
                        (j + 1 < got.size() && got[j + 1].distance == got[j].distance) || j + 1 == got.size();
      if (!tied) EXPECT_EQ(got[j].external_id, (*want)[j].external_id);
    }
  }
  uint64_t batches = 0, requests = 0;
  for (auto &sh : shards) {
    vkgpu_stats st{};
    EXPECT_EQ(vkgpu_get_stats(sh->handle(), &st), 0);
    batches += st.batches;
    requests += st.batched_requests;
  }
  EXPECT_EQ(requests, (uint64_t)kQueries * kShards);
  EXPECT_TRUE(batches < requests);  // the 16 concurrent requests of a shard did not run one by one
  std::fprintf(stderr, "  remote-mode fan-out: %llu requests in %llu batches over %d shards\n",
               (unsigned long long)requests, (unsigned long long)batches, kShards);
}

// In-memory chunk streams (the module's are RDBChunkOutputStream / RDBChunkInputStream)
struct MemoryStream : public OutputStream, public InputStream {
  std::vector<std::string> chunks;
  size_t next = 0;
  vks::Status SaveChunk(const char *data, size_t len) override {
    chunks.emplace_back(data, len);
    return vks::OkStatus();
  }
  vks::StatusOr<std::unique_ptr<std::string>> LoadChunk() override {
    if (next >= chunks.size()) return vks::NotFoundError("no more chunks");
    return std::make_unique<std::string>(chunks[next++]);
  }
  bool HasNext() const override { return next < chunks.size(); }
};

// SaveAndLoadFlat (testing/vector_test.cc:620-694): 50 queries answer the same (key and float distance) after a
// save + load; here additionally after swap-deletes, for COSINE (stored normalised, magnitudes in the key metadata)
// and with the byte layout of the element chunks checked.
static void SaveAndLoadFlat() {
  for (auto metric : {DistanceMetric::kL2, DistanceMetric::kCosine}) {
    auto r = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, metric, kInitialCap, kBlockSize));
    EXPECT_OK(r);
    if (!r.ok()) return;
    auto index = *r;
    auto vectors = DeterministicallyGenerateVectors(1000, kDimensions, 10.0);
    for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index.get(), vectors, i, ExpectedResults::kSuccess);
    for (int i = 3; i < 1000; i += 17) EXPECT_OK(index->RemoveRecord(IndexToKey(i)));  // swap-deletes reorder the slots
    MemoryStream data, keys;
    EXPECT_OK(index->SaveIndex(data));
    EXPECT_OK(index->SaveTrackedKeys(keys));
    const size_t live = index->GetTrackedKeyCount();
    EXPECT_EQ(data.chunks.size(), live + 1);
    EXPECT_EQ(keys.chunks.size(), live);
    BruteForceIndexHeader header;
    EXPECT_TRUE(header.ParseFromString(data.chunks[0]));
    EXPECT_EQ(header.curr_element_count, (uint64_t)live);
    EXPECT_EQ(header.size_per_element, (uint64_t)(kDimensions * 4 + 8));
    EXPECT_EQ(header.max_elements, (uint64_t)index->GetCapacity());
    for (size_t i = 1; i < data.chunks.size(); i++) EXPECT_EQ(data.chunks[i].size(), (size_t)(kDimensions * 4 + 8));
    // element chunk = the stored row + its label: slot 0 still holds key 0 (never moved)
    {
      uint64_t label0;
      std::memcpy(&label0, data.chunks[1].data() + kDimensions * 4, 8);
      auto key0 = index->GetKeyDuringSearch(label0);
      EXPECT_OK(key0);
      if (key0.ok()) EXPECT_EQ(*key0, IndexToKey(0));
    }
    auto loaded = VectorFlat<float>::LoadFromStream(CreateFlatVectorIndexProto(kDimensions, metric, kInitialCap, kBlockSize), data);
    EXPECT_OK(loaded);
    if (!loaded.ok()) continue;
    EXPECT_OK((*loaded)->LoadTrackedKeys(keys));
    EXPECT_EQ((*loaded)->GetTrackedKeyCount(), live);
    auto search_vectors = DeterministicallyGenerateVectors(50, kDimensions, 1.5);
    for (const auto &q : search_vectors) {
      auto a = index->Search(VectorToStr(q), 10, CancelNever());
      auto b = (*loaded)->Search(VectorToStr(q), 10, CancelNever());
      EXPECT_OK(a);
      EXPECT_OK(b);
      if (!a.ok() || !b.ok()) continue;
      EXPECT_EQ(a->size(), b->size());
      for (size_t j = 0; j < a->size() && j < b->size(); j++) {
        EXPECT_EQ((*a)[j].external_id, (*b)[j].external_id);
        EXPECT_TRUE(std::memcmp(&(*a)[j].distance, &(*b)[j].distance, 4) == 0);
      }
    }
    // a second save of the loaded index is byte-identical (slot order and rows survived)
    MemoryStream again;
    EXPECT_OK((*loaded)->SaveIndex(again));
    EXPECT_EQ(again.chunks.size(), data.chunks.size());
    for (size_t i = 1; i < again.chunks.size() && i < data.chunks.size(); i++) EXPECT_TRUE(again.chunks[i] == data.chunks[i]);
    // new keys continue after the largest loaded id (vector_base.cc:480-481)
    auto extra = DeterministicallyGenerateVectors(1, kDimensions, 3.0);
    EXPECT_OK((*loaded)->AddRecord("fresh_key", VectorToStr(extra[0])));
    // wrong dimension in the persisted header is rejected (bruteforce.h:190-193)
    MemoryStream bad = data;
    bad.next = 0;
    auto rejected = VectorFlat<float>::LoadFromStream(CreateFlatVectorIndexProto(kDimensions + 1, metric, kInitialCap, kBlockSize), bad);
    EXPECT_FALSE(rejected.ok());
  }
}

// SaveAndLoadHnsw (testing/vector_test.cc:502-580): save an EMPTY index, load it, populate it, recall >= 0.96, save,
// load again, recall >= 0.96.  Here additionally: the reloaded index answers every query exactly like the one that
// was saved (same graph on the same kernels), tombstones survive, and saving the reloaded index reproduces the
// stream byte for byte.
static void SaveAndLoadHnsw() {
  for (auto metric : {DistanceMetric::kCosine, DistanceMetric::kL2}) {
    const int initial_cap = 1000;
    const uint64_t k = 10;
    auto vectors = DeterministicallyGenerateVectors(1000, kDimensions, 2.2);
    auto index_flat = VectorFlat<float>::Create(CreateFlatVectorIndexProto(kDimensions, metric, initial_cap, kBlockSize));
    EXPECT_OK(index_flat);
    if (!index_flat.ok()) return;
    for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index_flat->get(), vectors, i, ExpectedResults::kSuccess);
    const VectorIndexProto hnsw_proto =
        CreateHNSWVectorIndexProto(kDimensions, metric, initial_cap, kM, kEFConstruction, kEFRuntime);
    MemoryStream empty_data, empty_keys;
    {
      auto index_hnsw = VectorHNSW<float>::Create(hnsw_proto);
      EXPECT_OK(index_hnsw);
      if (!index_hnsw.ok()) return;
      if (metric == DistanceMetric::kCosine) EXPECT_TRUE((*index_hnsw)->GetNormalize());
      EXPECT_OK((*index_hnsw)->SaveIndex(empty_data));
      EXPECT_OK((*index_hnsw)->SaveTrackedKeys(empty_keys));
      EXPECT_EQ(empty_data.chunks.size(), (size_t)1);  // header only (hnswalg.h:831-833)
      HNSWIndexHeader h;
      EXPECT_TRUE(h.ParseFromString(empty_data.chunks[0]));
      EXPECT_EQ(h.max_level, -1);
      EXPECT_EQ(h.enterpoint_node, 0xffffffffu);
      EXPECT_EQ(h.m, (uint64_t)kM);
      EXPECT_EQ(h.serialize_size_data_per_element, (uint64_t)(2 * kM * 4 + 4 + kDimensions * 4 + 8));
    }
    MemoryStream data, keys;
    std::vector<std::vector<Neighbor>> before;
    auto search_vectors = DeterministicallyGenerateVectors(50, kDimensions, 1.5);
    {
      auto loaded = VectorHNSW<float>::LoadFromStream(hnsw_proto, empty_data);
      EXPECT_OK(loaded);
      if (!loaded.ok()) return;
      EXPECT_OK((*loaded)->LoadTrackedKeys(empty_keys));
      for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(loaded->get(), vectors, i, ExpectedResults::kSuccess);
      for (int i = 5; i < 1000; i += 23) {  // tombstones in the graph, swap-deletes in the exact index
        EXPECT_OK((*loaded)->RemoveRecord(IndexToKey(i)));
        EXPECT_OK((*index_flat)->RemoveRecord(IndexToKey(i)));
      }
      EXPECT_TRUE(CalcRecall(index_flat->get(), loaded->get(), k, kDimensions, kEFRuntime) >= 0.96f);
      for (const auto &q : search_vectors) {
        auto r = (*loaded)->Search(VectorToStr(q), k, CancelNever());
        EXPECT_OK(r);
        before.push_back(r.ok() ? *r : std::vector<Neighbor>());
      }
      EXPECT_OK((*loaded)->SaveIndex(data));
      EXPECT_OK((*loaded)->SaveTrackedKeys(keys));
    }
    auto reloaded = VectorHNSW<float>::LoadFromStream(hnsw_proto, data);
    EXPECT_OK(reloaded);
    if (!reloaded.ok()) {
      std::fprintf(stderr, "  load: %s\n", reloaded.status().message().c_str());
      continue;
    }
    EXPECT_OK((*reloaded)->LoadTrackedKeys(keys));
    EXPECT_TRUE(CalcRecall(index_flat->get(), reloaded->get(), k, kDimensions, kEFRuntime) >= 0.96f);
    for (size_t qi = 0; qi < search_vectors.size(); qi++) {
      auto r = (*reloaded)->Search(VectorToStr(search_vectors[qi]), k, CancelNever());
      EXPECT_OK(r);
      if (!r.ok()) continue;
      EXPECT_EQ(r->size(), before[qi].size());
      for (size_t j = 0; j < r->size() && j < before[qi].size(); j++) {
        EXPECT_EQ((*r)[j].external_id, before[qi][j].external_id);
        EXPECT_TRUE(std::memcmp(&(*r)[j].distance, &before[qi][j].distance, 4) == 0);
      }
      for (const auto &nb : *r) {  // a tombstoned key is never returned
        const int id = std::atoi(nb.external_id.c_str());
        EXPECT_TRUE(!(id >= 5 && (id - 5) % 23 == 0));
      }
    }
    MemoryStream again;
    EXPECT_OK((*reloaded)->SaveIndex(again));
    EXPECT_EQ(again.chunks.size(), data.chunks.size());
    size_t differing = 0;
    for (size_t i = 1; i < again.chunks.size() && i < data.chunks.size(); i++) differing += again.chunks[i] != data.chunks[i];
    EXPECT_EQ(differing, (size_t)0);
    // the index stays usable after a load: new keys continue after the largest loaded id and are found
    // (alternating signs: no stored vector points the same way, so the COSINE answer is unique too)
    std::vector<std::vector<float>> extra(1, std::vector<float>(kDimensions));
    for (int j = 0; j < kDimensions; j++) extra[0][j] = (j % 2 ? -1.0f : 1.0f) * (float)(j + 1);
    EXPECT_OK((*reloaded)->AddRecord("fresh_key", VectorToStr(extra[0])));
    auto r = (*reloaded)->Search(VectorToStr(extra[0]), 1, CancelNever());
    EXPECT_OK(r);
    if (r.ok() && !r->empty()) EXPECT_EQ((*r)[0].external_id, std::string("fresh_key"));
    // a corrupted stream is rejected with the reference's message (hnswalg.h:1000)
    MemoryStream bad = data;
    bad.next = 0;
    uint16_t too_many = 2 * kM + 1;
    std::memcpy(bad.chunks[1].data(), &too_many, 2);
    auto rejected = VectorHNSW<float>::LoadFromStream(hnsw_proto, bad);
    EXPECT_FALSE(rejected.ok());
    EXPECT_TRUE(rejected.status().message().find("level-0 neighbor count exceeds 2*M") != std::string::npos);
  }
}

// vkgpu_stats.hops / distance_evals describe the MOST RECENT search call (bench.py divides them by that call's
// batch to state algorithmic bytes): repeating a call repeats the numbers, a batch is the sum of its queries.
static void HnswCountersPerCall() {
  auto index = VectorHNSW<float>::Create(CreateHNSWVectorIndexProto(kDimensions, DistanceMetric::kL2, 2000, kM, 100, 64));
  EXPECT_OK(index);
  if (!index.ok()) return;
  auto vectors = DeterministicallyGenerateVectors(2000, kDimensions, 3.0);
  for (size_t i = 0; i < vectors.size(); ++i) VerifyAdd(index->get(), vectors, i, ExpectedResults::kSuccess);
  auto queries = DeterministicallyGenerateVectors(8, kDimensions, 1.7);
  uint64_t hops[8], evals[8], sum_h = 0, sum_e = 0;
  for (int rep = 0; rep < 2; rep++)
    for (int i = 0; i < 8; i++) {
      EXPECT_OK((*index)->Search(VectorToStr(queries[i]), 10, CancelNever()));
      const vkgpu_stats st = (*index)->Stats();
      EXPECT_TRUE(st.hops > 0 && st.distance_evals >= st.hops);
      EXPECT_TRUE(st.distance_evals < 2000 + 512);  // level 0 evaluates a node at most once; the descent adds a few
      if (rep == 0) {
        hops[i] = st.hops;
        evals[i] = st.distance_evals;
        sum_h += st.hops;
        sum_e += st.distance_evals;
      } else {
        EXPECT_EQ(st.hops, hops[i]);
        EXPECT_EQ(st.distance_evals, evals[i]);
      }
    }
  std::string flat;
  for (const auto &q : queries) flat.append(reinterpret_cast<const char *>(q.data()), q.size() * 4);
  for (int rep = 0; rep < 2; rep++) {
    EXPECT_OK((*index)->SearchBatch(flat, 8, 10));
    const vkgpu_stats st = (*index)->Stats();
    EXPECT_EQ(st.hops, sum_h);
    EXPECT_EQ(st.distance_evals, sum_e);
  }
}

// ---- host-only modes for tests/test_hnsw_serialization.py: chunk streams cross the process boundary as one file,
// u64 chunk count, then per chunk u64 length + bytes (the same container oracle/ref_capi.cc uses).
static bool ReadStreamFile(const char *path, MemoryStream &out) {
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
    out.chunks.push_back(std::move(c));
  }
  std::fclose(f);
  return ok;
}
static bool WriteStreamFile(const char *path, const MemoryStream &s) {
  FILE *f = std::fopen(path, "wb");
  if (!f) return false;
  const uint64_t n = s.chunks.size();
  std::fwrite(&n, 8, 1, f);
  for (const auto &c : s.chunks) {
    const uint64_t len = c.size();
    std::fwrite(&len, 8, 1, f);
    std::fwrite(c.data(), 1, len, f);
  }
  return std::fclose(f) == 0;
}
static bool ReadFloats(const char *path, std::vector<float> &out) {
  FILE *f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize((size_t)bytes / 4);
  const bool ok = std::fread(out.data(), 4, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}
// --normalize IN DIM OUT: NormalizeEmbedding on every row of a raw float32 file; OUT = the normalised rows followed by
// one magnitude per row.  Host only (tests/test_host_cpp.py compares the bits with the C oracle and the Python mirror).
static int NormalizeMode(int argc, char **argv) {
  if (argc < 5) return 2;
  std::vector<float> X;
  if (!ReadFloats(argv[2], X)) return 2;
  const size_t dim = std::strtoull(argv[3], nullptr, 10), n = X.size() / dim;
  std::vector<float> out(n * dim), mags(n);
  for (size_t i = 0; i < n; i++) {
    auto r = NormalizeEmbedding(std::string_view(reinterpret_cast<const char *>(&X[i * dim]), dim * 4), sizeof(float), &mags[i]);
    std::memcpy(&out[i * dim], r.data(), dim * 4);
  }
  FILE *f = std::fopen(argv[4], "wb");
  if (!f) return 3;
  std::fwrite(out.data(), 4, out.size(), f);
  std::fwrite(mags.data(), 4, mags.size(), f);
  return std::fclose(f) == 0 ? 0 : 3;
}

// --flat-resave IN DIM [OUT]: LoadFlatHeader + LoadFlatElements on the stream; prints "OK count capacity" or
// "ERR message"; with OUT, SaveFlatImage of what was loaded goes there.  Host only.
static int FlatResaveMode(int argc, char **argv) {
  if (argc < 4) return 2;
  MemoryStream in;
  if (!ReadStreamFile(argv[2], in)) return 2;
  const size_t dim = std::strtoull(argv[3], nullptr, 10);
  auto header = LoadFlatHeader(in, dim);
  if (!header.ok()) {
    std::printf("ERR %s\n", header.status().message().c_str());
    return 0;
  }
  std::vector<float> rows;
  std::vector<uint64_t> labels;
  const vks::Status s = LoadFlatElements(in, *header, dim, [&](const uint64_t *l, const float *r, uint64_t n) {
    labels.insert(labels.end(), l, l + n);
    rows.insert(rows.end(), r, r + n * dim);
    return vks::OkStatus();
  });
  if (!s.ok()) {
    std::printf("ERR %s\n", s.message().c_str());
    return 0;
  }
  std::printf("OK %llu %llu\n", (unsigned long long)labels.size(), (unsigned long long)header->max_elements);
  if (argc > 4) {
    MemoryStream out;
    const vks::Status w = SaveFlatImage(labels.size(), header->max_elements, dim,
                                        [&](uint64_t first, uint64_t n, float *r, uint64_t *l) {
                                          std::memcpy(r, rows.data() + first * dim, n * dim * sizeof(float));
                                          std::memcpy(l, labels.data() + first, n * sizeof(uint64_t));
                                          return vks::OkStatus();
                                        },
                                        out);
    if (!w.ok() || !WriteStreamFile(argv[4], out)) return 3;
  }
  return 0;
}

// --hnsw-load IN DIM CAP M VALIDATE [OUT]: LoadHnswImage on the stream; prints "OK n max_level enterpoint
// max_elements duplicates deleted" or "ERR message"; with OUT, SaveHnswImage of what was loaded goes there.
static int HnswLoadMode(int argc, char **argv) {
  if (argc < 7) return 2;
  MemoryStream in;
  if (!ReadStreamFile(argv[2], in)) return 2;
  const size_t dim = std::strtoull(argv[3], nullptr, 10), cap = std::strtoull(argv[4], nullptr, 10);
  const size_t m = std::strtoull(argv[5], nullptr, 10);
  const bool validate = std::atoi(argv[6]) != 0;
  auto loaded = LoadHnswImage(in, dim, cap, m, validate);
  if (!loaded.ok()) {
    std::printf("ERR %s\n", loaded.status().message().c_str());
    return 0;
  }
  const HnswGraphImage &g = loaded->image;
  uint64_t deleted = 0;
  for (uint8_t d : g.deleted) deleted += d;
  std::printf("OK %llu %d %u %llu %llu %llu\n", (unsigned long long)g.n, g.max_level, g.enterpoint,
              (unsigned long long)loaded->max_elements, (unsigned long long)loaded->duplicate_labels,
              (unsigned long long)deleted);
  if (argc > 7) {
    MemoryStream out;
    const HnswRowFetcher rows = [&](uint64_t first, uint64_t count, float *dst) {
      std::memcpy(dst, g.vecs.data() + first * dim, count * dim * sizeof(float));
      return vks::OkStatus();
    };
    // max_elements as the header recorded it: the reference re-saves max_elements_ = the resolved capacity
    const vks::Status s = SaveHnswImage(g, dim, loaded->max_elements, loaded->ef_construction, rows, out);
    if (!s.ok() || !WriteStreamFile(argv[7], out)) return 3;
  }
  return 0;
}

// ---- GPU modes for tests/test_hnsw_interchange_gpu.py (files cross between the reference and the GPU index)
// keys of a loaded stream: one TrackedKeyMetadata chunk per level-0 record, key = decimal label
static MemoryStream KeysOf(const MemoryStream &data, size_t dim, size_t m) {
  MemoryStream keys;
  HNSWIndexHeader h;
  if (!h.ParseFromString(data.chunks[0])) return keys;
  const size_t label_off = 2 * m * 4 + 4 + dim * 4;
  for (uint64_t i = 0; i < h.curr_element_count; i++) {
    uint64_t label;
    std::memcpy(&label, data.chunks[1 + i].data() + label_off, 8);
    uint32_t word;
    std::memcpy(&word, data.chunks[1 + i].data(), 4);
    if (word & (1u << 16)) continue;  // tombstoned elements have no key any more
    TrackedKeyMetadataPb pb;
    pb.key = std::to_string(label);
    pb.internal_id = label;
    pb.magnitude = kDefaultMagnitude;
    const std::string b = pb.SerializeAsString();
    keys.chunks.emplace_back(b);
  }
  return keys;
}
// --hnsw-gpu-load IN DIM CAP M QUERIES K EF RESULTS RESAVED: a stream written by the CPU module -> LoadFromStream ->
// search every query on the GPU (results: per query u32 n, then n x {u64 label, f32 distance}) -> SaveIndex again.
static int HnswGpuLoadMode(int argc, char **argv) {
  if (argc < 11) return 2;
  MemoryStream in;
  if (!ReadStreamFile(argv[2], in)) return 2;
  const int dim = std::atoi(argv[3]), cap = std::atoi(argv[4]), m = std::atoi(argv[5]);
  std::vector<float> Q;
  if (!ReadFloats(argv[6], Q)) return 2;
  const uint64_t k = std::strtoull(argv[7], nullptr, 10);
  const size_t ef = std::strtoull(argv[8], nullptr, 10);
  auto loaded = VectorHNSW<float>::LoadFromStream(CreateHNSWVectorIndexProto(dim, DistanceMetric::kL2, cap, m, 200, 10), in);
  if (!loaded.ok()) {
    std::printf("ERR %s\n", loaded.status().message().c_str());
    return 1;
  }
  MemoryStream keys = KeysOf(in, dim, m);
  if (!(*loaded)->LoadTrackedKeys(keys).ok()) return 3;
  FILE *f = std::fopen(argv[9], "wb");
  if (!f) return 3;
  for (size_t q = 0; q + dim <= Q.size(); q += dim) {
    auto r = (*loaded)->Search(std::string_view(reinterpret_cast<const char *>(&Q[q]), (size_t)dim * 4), k, CancelNever(),
                               nullptr, ef);
    if (!r.ok()) {
      std::printf("ERR %s\n", r.status().message().c_str());
      return 1;
    }
    const uint32_t n = (uint32_t)r->size();
    std::fwrite(&n, 4, 1, f);
    for (const auto &nb : *r) {
      const uint64_t label = std::strtoull(nb.external_id.c_str(), nullptr, 10);
      std::fwrite(&label, 8, 1, f);
      std::fwrite(&nb.distance, 4, 1, f);
    }
  }
  std::fclose(f);
  MemoryStream out;
  if (!(*loaded)->SaveIndex(out).ok() || !WriteStreamFile(argv[10], out)) return 3;
  std::printf("OK %zu\n", (*loaded)->GetTrackedKeyCount());
  return 0;
}
// --hnsw-gpu-build VECTORS DIM M EFC DELETE_EVERY OUT: build on the GPU through AddRecord (keys = decimal row index),
// tombstone every DELETE_EVERY-th key, SaveIndex -> OUT (for the CPU module to load).
static int HnswGpuBuildMode(int argc, char **argv) {
  if (argc < 8) return 2;
  std::vector<float> X;
  if (!ReadFloats(argv[2], X)) return 2;
  const int dim = std::atoi(argv[3]), m = std::atoi(argv[4]), efc = std::atoi(argv[5]), every = std::atoi(argv[6]);
  const size_t n = X.size() / dim;
  auto index = VectorHNSW<float>::Create(CreateHNSWVectorIndexProto(dim, DistanceMetric::kL2, (int)n, m, efc, 10));
  if (!index.ok()) {
    std::printf("ERR %s\n", index.status().message().c_str());
    return 1;
  }
  for (size_t i = 0; i < n; i++) {
    auto r = (*index)->AddRecord(std::to_string(i), std::string_view(reinterpret_cast<const char *>(&X[i * dim]), (size_t)dim * 4));
    if (!r.ok() || *r != RecordResult::kAdded) return 3;
  }
  if (every > 0)
    for (size_t i = 1; i < n; i += every)
      if (!(*index)->RemoveRecord(std::to_string(i)).ok()) return 3;
  MemoryStream out;
  if (!(*index)->SaveIndex(out).ok() || !WriteStreamFile(argv[7], out)) return 3;
  std::printf("OK %zu\n", (*index)->GetTrackedKeyCount());
  return 0;
}

// --hnsw-perf N DIM BATCH K EF REPS: clustered corpus built on the GPU with one vkgpu_add_batch, then the same batch
// searched REPS times per setting of the environment variable named in VAR (unset / set), wall-clock per batch.
// A quick A/B of a kernel switch inside one process; the numbers that count are bench.py's.
static int HnswPerfMode(int argc, char **argv) {
  if (argc < 9) return 2;
  const size_t n = std::strtoull(argv[2], nullptr, 10);
  const int dim = std::atoi(argv[3]), batch = std::atoi(argv[4]), k = std::atoi(argv[5]), ef = std::atoi(argv[6]);
  const int reps = std::atoi(argv[7]);
  const char *var = argv[8];
  uint64_t rng = 88172645463325252ull;
  auto uni = [&]() {
    rng ^= rng << 13;
    rng ^= rng >> 7;
    rng ^= rng << 17;
    return (float)((rng >> 11) * (1.0 / 9007199254740992.0));
  };
  auto gauss = [&]() {  // Irwin-Hall, good enough for a timing corpus
    float s = 0;
    for (int i = 0; i < 12; i++) s += uni();
    return s - 6.0f;
  };
  const size_t ncent = std::max<size_t>(n / 1000, 8);
  std::vector<float> cent(ncent * dim), X(n * dim), Q((size_t)batch * dim);
  for (auto &v : cent) v = gauss();
  for (size_t i = 0; i < n; i++) {
    const size_t c = (size_t)(uni() * ncent) % ncent;
    for (int j = 0; j < dim; j++) X[i * dim + j] = cent[c * dim + j] + 0.3f * gauss();
  }
  for (int i = 0; i < batch; i++) {
    const size_t c = (size_t)(uni() * ncent) % ncent;
    for (int j = 0; j < dim; j++) Q[(size_t)i * dim + j] = cent[c * dim + j] + 0.3f * gauss();
  }
  auto index = VectorHNSW<float>::Create(CreateHNSWVectorIndexProto(dim, DistanceMetric::kL2, (int)n, 16, 200, ef));
  if (!index.ok()) {
    std::printf("ERR %s\n", index.status().message().c_str());
    return 1;
  }
  std::vector<uint64_t> labels(n);
  for (size_t i = 0; i < n; i++) labels[i] = i;
  timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (vkgpu_add_batch((*index)->handle(), labels.data(), X.data(), n) != 0) return 3;
  clock_gettime(CLOCK_MONOTONIC, &t1);
  std::printf("build %zu x %d: %.2f s\n", n, dim, (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9);
  std::vector<float> od((size_t)batch * k), od2((size_t)batch * k);
  std::vector<uint64_t> ol((size_t)batch * k), ol2((size_t)batch * k);
  std::vector<uint32_t> on(batch);
  for (int setting = 0; setting < 4; setting++) {
    const bool set = setting & 1;
    // VAR or VAR=VALUE: the second setting of each pair exports VALUE (default "1"), the first leaves VAR unset
    const std::string vs(var);
    const size_t eq = vs.find('=');
    const std::string vname = vs.substr(0, eq), vval = eq == std::string::npos ? "1" : vs.substr(eq + 1);
    if (set) setenv(vname.c_str(), vval.c_str(), 1); else unsetenv(vname.c_str());
    auto &d = set ? od2 : od;
    auto &l = set ? ol2 : ol;
    for (int w = 0; w < 3; w++)
      if (vkgpu_search_batch((*index)->handle(), Q.data(), batch, k, ef, nullptr, 0, d.data(), l.data(), on.data()) != 0) return 3;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++)
      if (vkgpu_search_batch((*index)->handle(), Q.data(), batch, k, ef, nullptr, 0, d.data(), l.data(), on.data()) != 0) return 3;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms = ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6) / reps;
    const vkgpu_stats st = (*index)->Stats();
    std::printf("%s=%s: %.3f ms per batch of %d (host buffers), %.0f QPS; hops/query %.1f, evaluations/query %.1f\n", var,
                set ? "1" : "unset", ms, batch, batch / ms * 1e3, (double)st.hops / batch, (double)st.distance_evals / batch);
  }
  std::printf("results identical across settings: %s\n",
              (od == od2 && ol == ol2) ? "yes" : "NO");
  return (od == od2 && ol == ol2) ? 0 : 1;
}

static std::string Hex(const std::string &s) {
  static const char *d = "0123456789abcdef";
  std::string out;
  for (unsigned char c : s) {
    out.push_back(d[c >> 4]);
    out.push_back(d[c & 15]);
  }
  return out;
}
// `--wire`: prints the hand-encoded protobuf messages for a fixed list of cases; tests/test_host_cpp.py compares
// them with what the protobuf runtime produces for the reference's message definitions.
static int PrintWire() {
  const uint64_t hdr[][3] = {{15000, 408, 100}, {0, 0, 0}, {10240, 3080, 0}, {1ull << 40, 12, 300}, {1, 1, 1}};
  for (const auto &h : hdr) {
    BruteForceIndexHeader m;
    m.max_elements = h[0];
    m.size_per_element = h[1];
    m.curr_element_count = h[2];
    BruteForceIndexHeader back;
    if (!back.ParseFromString(m.SerializeAsString()) || back.max_elements != h[0] || back.size_per_element != h[1] ||
        back.curr_element_count != h[2])
      return 2;
    std::printf("header %llu %llu %llu %s\n", (unsigned long long)h[0], (unsigned long long)h[1], (unsigned long long)h[2],
                Hex(m.SerializeAsString()).c_str());
  }
  {
    struct {
      uint64_t cap, count;
      int32_t max_level;
      uint32_t ep;
      uint64_t m;
      double mult;
      uint64_t efc;
    } cases[] = {{1000, 0, -1, 0xffffffffu, 16, 1 / std::log(16.0), 20},
                 {32, 8, 2, 0, 16, 1 / std::log(16.0), 200},
                 {10240000, 10000000, 5, 123456, 32, 1 / std::log(32.0), 400},
                 {0, 0, 0, 0, 0, 0.0, 0},
                 {5, 5, 0, 4, 2, -0.0, 1}};
    for (const auto &c : cases) {
      HNSWIndexHeader m;
      m.max_elements = c.cap;
      m.curr_element_count = c.count;
      m.serialize_size_data_per_element = c.m * 8 + 4 + 400 + 8;
      m.label_offset = ((c.m * 8 + 4 + 7) & ~7ull) + 8;
      m.offset_data = c.m * 8 + 4;
      m.max_level = c.max_level;
      m.enterpoint_node = c.ep;
      m.max_m = c.m;
      m.max_m_0 = 2 * c.m;
      m.m = c.m;
      m.mult = c.mult;
      m.ef_construction = c.efc;
      HNSWIndexHeader back;
      if (!back.ParseFromString(m.SerializeAsString()) || back.max_level != c.max_level || back.enterpoint_node != c.ep ||
          std::memcmp(&back.mult, &c.mult, 8) != 0 || back.ef_construction != c.efc || back.max_m_0 != 2 * c.m)
        return 2;
      uint64_t bits;
      std::memcpy(&bits, &c.mult, 8);
      std::printf("hnsw %llu %llu %d %u %llu %016llx %llu %s\n", (unsigned long long)c.cap, (unsigned long long)c.count,
                  c.max_level, c.ep, (unsigned long long)c.m, (unsigned long long)bits, (unsigned long long)c.efc,
                  Hex(m.SerializeAsString()).c_str());
    }
  }
  struct {
    const char *key;
    uint64_t id;
    float mag;
  } keys[] = {{"0_key", 5, -1.0f}, {"", 0, 0.0f}, {"doc:1", 0, 1.5f}, {"k", 1ull << 33, 0.0f}, {"x", 7, -0.0f}};
  for (const auto &k : keys) {
    TrackedKeyMetadataPb m;
    m.key = k.key;
    m.internal_id = k.id;
    m.magnitude = k.mag;
    TrackedKeyMetadataPb back;
    if (!back.ParseFromString(m.SerializeAsString()) || back.key != k.key || back.internal_id != k.id ||
        std::memcmp(&back.magnitude, &k.mag, 4) != 0)
      return 2;
    uint32_t bits;
    std::memcpy(&bits, &k.mag, 4);
    std::printf("key %s|%llu|%08x %s\n", k.key, (unsigned long long)k.id, bits, Hex(m.SerializeAsString()).c_str());
  }
  return 0;
}

// host-only: normalisation arithmetic (vector_base.cc:112-138) and the no-CPU-fallback contract
// Cross-shard aggregation (src/query/fanout.cc:50-64, 153-212)
static void FanoutMerge() {
  using valkey_search::query::fanout::SearchPartitionResultsTracker;
  auto keys_of = [](const std::vector<Neighbor> &v) {
    std::string s;
    for (const auto &n : v) s += n.external_id + " ";
    return s;
  };
  {  // distinct distances: the merge of the shards' top-k is the global top-k, ascending, whatever the arrival order
    std::vector<Neighbor> all;
    for (int i = 0; i < 40; i++) all.emplace_back("k" + std::to_string(i), (float)((i * 37) % 41));
    std::vector<Neighbor> sorted = all;
    std::sort(sorted.begin(), sorted.end(), [](const Neighbor &a, const Neighbor &b) { return a.distance < b.distance; });
    for (int order = 0; order < 2; order++) {
      SearchPartitionResultsTracker tracker(7);
      for (int shard = 0; shard < 4; shard++) {
        const int sh = order ? 3 - shard : shard;
        std::vector<Neighbor> local;
        for (int i = sh; i < 40; i += 4) local.push_back(all[i]);
        std::sort(local.begin(), local.end(), [](const Neighbor &a, const Neighbor &b) { return a.distance < b.distance; });
        local.resize(7);  // each shard answers its own top-k
        tracker.AddResults(local);
      }
      auto merged = tracker.TakeNeighbors();
      EXPECT_EQ(merged.size(), (size_t)7);
      for (size_t i = 0; i < merged.size(); i++) EXPECT_EQ(merged[i].external_id, sorted[i].external_id);
    }
  }
  {  // equal distances: descending key order inside the tie; once full, an equal distance is NOT admitted
    SearchPartitionResultsTracker tracker(3);
    std::vector<Neighbor> a = {{"b", 1.0f}, {"d", 1.0f}}, b = {{"a", 1.0f}, {"c", 1.0f}, {"e", 0.5f}};
    tracker.AddResults(a);
    tracker.AddResults(b);  // "a" fills the heap, "c" ties with the worst and stays out, "e" evicts the smallest key
    EXPECT_EQ(keys_of(tracker.TakeNeighbors()), std::string("e d b "));
    SearchPartitionResultsTracker other(3);
    a = {{"b", 1.0f}, {"d", 1.0f}};  // AddResults moves the neighbours out, as the reference does
    b = {{"a", 1.0f}, {"c", 1.0f}, {"e", 0.5f}};
    other.AddResults(b);
    other.AddResults(a);  // the other arrival order keeps a different tie: documented, not hidden
    EXPECT_EQ(keys_of(other.TakeNeighbors()), std::string("e c a "));
  }
}

static void HostOnly(bool have_gpu) {
  FanoutMerge();
  {  // NormalizeStringRecordTests (testing/vector_test.cc:303-351)
    struct {
      const char *record;
      bool success;
      std::vector<float> expected;
    } cases[] = {{"[ 0.1]", true, {0.1f}},
                 {"[,0.1]", true, {0.1f}},
                 {"[ 0.1, ,0.2,0.3,]", true, {0.1f, 0.2f, 0.3f}},
                 {"[ 0.1, ,0.2,a,]", false, {}},
                 {"1.5, -2e3", true, {1.5f, -2000.0f}},  // brackets are optional
                 {"", true, {}}};
    for (const auto &c : cases) {
      auto r = VectorBase::NormalizeStringRecord(c.record);
      EXPECT_EQ(r.has_value(), c.success);
      if (!r || !c.success) continue;
      EXPECT_EQ(r->size(), c.expected.size() * sizeof(float));
      for (size_t i = 0; i < c.expected.size() && (i + 1) * 4 <= r->size(); i++) {
        float value;
        std::memcpy(&value, r->data() + i * 4, 4);
        EXPECT_TRUE(value == c.expected[i]);
      }
    }
  }
  std::vector<float> v = {3.0f, 4.0f, 0.0f};
  float mag = 0;
  auto n = NormalizeEmbedding(VectorToStr(v), sizeof(float), &mag);
  const float *f = reinterpret_cast<const float *>(n.data());
  EXPECT_TRUE(mag == 5.0f);
  EXPECT_TRUE(f[0] == 0.2f * 3.0f && f[1] == 0.2f * 4.0f && f[2] == 0.0f);
  std::vector<float> z(8, 0.0f);
  auto nz = NormalizeEmbedding(VectorToStr(z), sizeof(float), &mag);
  EXPECT_TRUE(mag == 0.0f && reinterpret_cast<const float *>(nz.data())[3] == 0.0f);  // zero vector stays zero
  if (!have_gpu) {
    auto r = VectorFlat<float>::Create(CreateFlatVectorIndexProto(16, DistanceMetric::kL2, 100, 10));
    EXPECT_FALSE(r.ok());  // no device => error status, never a CPU index
    if (!r.ok()) std::fprintf(stderr, "  Create without a GPU: %s\n", r.status().message().c_str());
  }
}

int main(int argc, char **argv) {
  if (argc > 1 && std::string(argv[1]) == "--wire") return PrintWire();
  if (argc > 1 && std::string(argv[1]) == "--hnsw-load") return HnswLoadMode(argc, argv);
  if (argc > 1 && std::string(argv[1]) == "--flat-resave") return FlatResaveMode(argc, argv);
  if (argc > 1 && std::string(argv[1]) == "--normalize") return NormalizeMode(argc, argv);
  if (argc > 1 && std::string(argv[1]) == "--hnsw-gpu-load") return HnswGpuLoadMode(argc, argv);
  if (argc > 1 && std::string(argv[1]) == "--hnsw-gpu-build") return HnswGpuBuildMode(argc, argv);
  if (argc > 1 && std::string(argv[1]) == "--hnsw-perf") return HnswPerfMode(argc, argv);
  const bool host_only = argc > 1 && std::string(argv[1]) == "--host-only";
  struct Case {
    const char *name;
    void (*fn)();
  } cases[] = {{"BasicFlat", BasicFlat},
               {"BasicHNSW", BasicHNSW},
               {"EfRuntimeRecall", EfRuntimeRecall},
               {"IntegrationCosineGoldens", IntegrationCosineGoldens},
               {"Prefilter", Prefilter},
               {"SaveAndLoadFlat", SaveAndLoadFlat},
               {"SaveAndLoadHnsw", SaveAndLoadHnsw},
               {"HnswCountersPerCall", HnswCountersPerCall},
               {"InlineFilterAndBatch", InlineFilterAndBatch},
               {"RemoteModeFanout", RemoteModeFanout}};
  {
    const int before = g_failures;
    HostOnly(vkgpu_device_count() > 0);
    std::printf("[%s] HostOnly\n", g_failures == before ? "  OK  " : "FAILED");
  }
  const std::string only = argc > 2 && std::string(argv[1]) == "--case" ? argv[2] : "";
  if (!host_only)
    for (const auto &c : cases) {
      if (!only.empty() && only != c.name) continue;
      const int before = g_failures;
      c.fn();
      std::printf("[%s] %s\n", g_failures == before ? "  OK  " : "FAILED", c.name);
    }
  std::printf("%d checks, %d failures\n", g_checks, g_failures);
  return g_failures ? 1 : 0;
}
