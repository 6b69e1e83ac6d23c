// See hnsw_serialization.h.  Restates HierarchicalNSW::SaveIndex / LoadIndex (third_party/hnswlib/hnswalg.h:808-1139)
// over flat arrays; every validation message is the reference's.
#include "hnsw_serialization.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

namespace valkey_search::indexes {

namespace {

constexpr uint32_t kCountMask = 0xffffu;       // getListCount reads an unsigned short (hnswalg.h getListCount)
constexpr uint32_t kDeleteMark = 0x01u << 16;  // DELETE_MARK in the third byte of the level-0 count word
constexpr size_t kU32 = sizeof(uint32_t);
constexpr size_t kLabelBytes = sizeof(uint64_t);  // labeltype = size_t (hnswlib.h:141)

void PutVarint(std::string &out, uint64_t v) {
  while (v >= 0x80) {
    out.push_back((char)((v & 0x7f) | 0x80));
    v >>= 7;
  }
  out.push_back((char)v);
}
bool GetVarint(std::string_view s, size_t &pos, uint64_t &v) {
  v = 0;
  for (int shift = 0; shift < 64 && pos < s.size(); shift += 7) {
    const uint8_t b = (uint8_t)s[pos++];
    v |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}
void PutU64Field(std::string &out, int field, uint64_t v) {
  if (!v) return;  // proto3: defaults are not written
  PutVarint(out, (uint64_t)field << 3);
  PutVarint(out, v);
}

struct LoadFailure {
  std::string msg;
};

}  // namespace

std::string HNSWIndexHeader::SerializeAsString() const {
  std::string out;
  PutU64Field(out, 1, offset_level_0);
  PutU64Field(out, 2, max_elements);
  PutU64Field(out, 3, curr_element_count);
  PutU64Field(out, 4, serialize_size_data_per_element);
  PutU64Field(out, 5, label_offset);
  PutU64Field(out, 6, offset_data);
  PutU64Field(out, 7, (uint64_t)(int64_t)max_level);  // int32: sign-extended to 64 bits on the wire
  PutU64Field(out, 8, enterpoint_node);
  PutU64Field(out, 9, max_m);
  PutU64Field(out, 10, max_m_0);
  PutU64Field(out, 11, m);
  uint64_t bits;
  std::memcpy(&bits, &mult, 8);
  if (bits) {
    PutVarint(out, (12u << 3) | 1);
    out.append(reinterpret_cast<const char *>(&bits), 8);  // little endian, as on every host this runs on
  }
  PutU64Field(out, 13, ef_construction);
  return out;
}

bool HNSWIndexHeader::ParseFromString(std::string_view s) {
  *this = HNSWIndexHeader();
  size_t pos = 0;
  while (pos < s.size()) {
    uint64_t tag, v = 0;
    if (!GetVarint(s, pos, tag)) return false;
    const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
    if (wt == 0) {
      if (!GetVarint(s, pos, v)) return false;
      switch (field) {
        case 1: offset_level_0 = v; break;
        case 2: max_elements = v; break;
        case 3: curr_element_count = v; break;
        case 4: serialize_size_data_per_element = v; break;
        case 5: label_offset = v; break;
        case 6: offset_data = v; break;
        case 7: max_level = (int32_t)v; break;
        case 8: enterpoint_node = (uint32_t)v; break;
        case 9: max_m = v; break;
        case 10: max_m_0 = v; break;
        case 11: m = v; break;
        case 13: ef_construction = v; break;
        default: break;
      }
    } else if (wt == 1) {
      if (pos + 8 > s.size()) return false;
      if (field == 12) std::memcpy(&mult, s.data() + pos, 8);
      pos += 8;
    } else if (wt == 2) {
      if (!GetVarint(s, pos, v) || pos + v > s.size()) return false;
      pos += v;
    } else if (wt == 5) {
      if (pos + 4 > s.size()) return false;
      pos += 4;
    } else {
      return false;
    }
  }
  return true;
}

Status SaveHnswImage(const HnswGraphImage &g, size_t dim, uint64_t max_elements, uint64_t ef_construction,
                     const HnswRowFetcher &rows, OutputStream &output) {
  const size_t maxM = g.M, maxM0 = 2 * (size_t)g.M;
  const size_t size_links_level0 = maxM0 * kU32 + kU32;                    // hnswalg.h:152
  const size_t offset_data = (size_links_level0 + 7) & ~(size_t)7;        // :153-154
  const size_t vector_size = dim * sizeof(float);
  const size_t element_bytes = size_links_level0 + vector_size + kLabelBytes;  // :156-157
  const size_t stride = maxM * kU32 + kU32;                                // size_links_per_element_, :174-175
  if (g.levels.size() != g.n || g.labels.size() != g.n || g.cnt0.size() != g.n || g.links0.size() != g.n * maxM0 ||
      (g.n && g.upper_offset.size() != g.n))
    return vks::InternalError("HNSW graph image is inconsistent");

  HNSWIndexHeader header;
  header.offset_level_0 = 0;
  header.max_elements = std::max<uint64_t>(max_elements, g.n);
  header.curr_element_count = g.n;
  header.serialize_size_data_per_element = element_bytes;
  header.label_offset = offset_data + sizeof(char *);
  header.offset_data = size_links_level0;  // the reference writes size_links_level0_ into this field (:815)
  header.max_level = g.max_level;
  header.enterpoint_node = g.enterpoint;
  header.max_m = maxM;
  header.max_m_0 = maxM0;
  header.m = g.M;
  header.mult = 1 / std::log(1.0 * g.M);  // :176
  header.ef_construction = ef_construction;
  const std::string serialized = header.SerializeAsString();
  VKS_RETURN_IF_ERROR(output.SaveChunk(serialized.data(), serialized.size()));
  if (g.n == 0) return vks::OkStatus();

  constexpr uint64_t kBlock = 4096;
  std::vector<float> block(kBlock * dim);
  std::vector<char> buf(element_bytes);
  for (uint64_t first = 0; first < g.n; first += kBlock) {
    const uint64_t count = std::min<uint64_t>(kBlock, g.n - first);
    VKS_RETURN_IF_ERROR(rows(first, count, block.data()));
    for (uint64_t j = 0; j < count; j++) {
      const uint64_t i = first + j;
      const uint32_t word = (g.cnt0[i] & kCountMask) | ((!g.deleted.empty() && g.deleted[i]) ? kDeleteMark : 0u);
      std::memcpy(buf.data(), &word, kU32);
      std::memcpy(buf.data() + kU32, &g.links0[i * maxM0], maxM0 * kU32);
      std::memcpy(buf.data() + size_links_level0, &block[j * dim], vector_size);
      std::memcpy(buf.data() + size_links_level0 + vector_size, &g.labels[i], kLabelBytes);
      VKS_RETURN_IF_ERROR(output.SaveChunk(buf.data(), buf.size()));
    }
  }
  std::vector<char> lists;
  for (uint64_t i = 0; i < g.n; i++) {
    const uint64_t level = g.levels[i] > 0 ? (uint64_t)g.levels[i] : 0;
    const uint64_t link_list_size = level * stride;  // little-endian u64 on every host this runs on (htole64, :851)
    VKS_RETURN_IF_ERROR(output.SaveChunk(reinterpret_cast<const char *>(&link_list_size), sizeof(uint64_t)));
    if (!link_list_size) continue;
    lists.assign(link_list_size, 0);
    for (uint64_t l = 0; l < level; l++) {
      const uint64_t b = g.upper_offset[i] + l;
      if (b >= g.upper_cnt.size() || (b + 1) * maxM > g.upper_links.size())
        return vks::InternalError("HNSW graph image is inconsistent");
      const uint32_t word = g.upper_cnt[b] & kCountMask;
      std::memcpy(lists.data() + l * stride, &word, kU32);
      std::memcpy(lists.data() + l * stride + kU32, &g.upper_links[b * maxM], maxM * kU32);
    }
    VKS_RETURN_IF_ERROR(output.SaveChunk(lists.data(), lists.size()));
  }
  return vks::OkStatus();
}

StatusOr<HnswLoadResult> LoadHnswImage(InputStream &input, size_t dim, size_t max_elements_i, size_t expected_m,
                                       bool validate) {
  // `soft` checks are the ones the kill switch bypasses; `hard` ones protect the GPU arrays and always apply
  auto fail = [](std::string_view msg) {
    throw LoadFailure{std::string("HNSWLib error while loading an index: HNSW index load validation failed: ") +
                      std::string(msg)};
  };
  auto hard = [&](bool ok, std::string_view msg) {
    if (!ok) fail(msg);
  };
  auto soft = [&](bool ok, std::string_view msg) {
    if (!ok && validate) fail(msg);
  };
  try {
    auto serialized_header = input.LoadChunk();
    if (!serialized_header.ok()) return serialized_header.status();
    HNSWIndexHeader header;
    if (!header.ParseFromString(**serialized_header)) return vks::InternalError("Could not deserialize HNSW header");

    HnswLoadResult out;
    HnswGraphImage &g = out.image;
    const uint64_t cur = header.curr_element_count;
    const size_t maxM = header.max_m, maxM0 = header.max_m_0, M = header.m;
    const size_t vector_size = dim * sizeof(float);
    const size_t size_links_level0 = maxM0 * kU32 + kU32;
    const size_t stride = maxM * kU32 + kU32;
    out.max_elements = std::max<uint64_t>(cur, std::max<uint64_t>(max_elements_i, header.max_elements));
    out.ef_construction = header.ef_construction;

    {  // header validation, in the reference's order (hnswalg.h:930-981)
      const size_t exp_m = expected_m > 10000 ? 10000 : expected_m;
      hard(exp_m >= 1, "M must be >= 1");
      soft(M == exp_m, "header M does not match index definition");
      hard(M >= 1 && M <= 10000, "header M does not match index definition");
      hard(maxM == M, "header maxM does not equal M");
      hard(maxM0 == 2 * M, "header maxM0 does not equal 2*M");
      hard(maxM0 <= 0xFFFF, "maxM0 exceeds the 16-bit neighbor-count field");
      hard(vector_size > 0, "vector size must be > 0");
      hard(header.serialize_size_data_per_element == size_links_level0 + vector_size + kLabelBytes,
           "serialized element size is inconsistent with the geometry");
      soft(header.offset_level_0 == 0, "offset_level_0 must be 0");
      if (M >= 2) {
        const double expected_mult = 1.0 / std::log((double)M);
        soft(header.mult > 0.0 && std::fabs(header.mult - expected_mult) <= 1e-6 * expected_mult,
             "mult is inconsistent with M");
      }
      hard(cur <= out.max_elements, "curr_element_count exceeds max_elements");
      if (cur == 0) {
        soft(header.max_level == -1 || header.max_level == 0, "empty index has a non-trivial max_level");
      } else {
        hard(header.max_level >= 0, "non-empty index has a negative max_level");
        soft((int64_t)header.max_level <= (int64_t)cur, "max_level exceeds the element count");
        hard(header.enterpoint_node < cur, "enterpoint_node is out of range");
      }
      hard(cur < 0xffffffffull, "level-0 allocation size overflows");  // tableint ids
    }

    g.M = (uint32_t)M;
    g.n = cur;
    g.max_level = cur ? header.max_level : -1;
    g.enterpoint = cur ? header.enterpoint_node : 0xffffffffu;
    // The arrays grow with the chunks that actually arrive: an absurd element count in a corrupt header runs out of
    // stream long before it runs out of memory.
    const uint64_t reserve = std::min<uint64_t>(cur, 1u << 20);
    g.levels.reserve(reserve);
    g.labels.reserve(reserve);
    g.deleted.reserve(reserve);
    g.cnt0.reserve(reserve);
    g.upper_offset.reserve(reserve);
    g.links0.reserve(reserve * maxM0);
    g.vecs.reserve(reserve * dim);

    // level-0 records (hnswalg.h:986-1015)
    for (uint64_t i = 0; i < cur; i++) {
      auto chunk = input.LoadChunk();
      if (!chunk.ok()) return chunk.status();
      const std::string &c = **chunk;
      hard(c.size() == size_links_level0 + vector_size + kLabelBytes, "level-0 element chunk has the wrong size");
      g.levels.push_back(0);
      g.labels.push_back(0);
      g.deleted.push_back(0);
      g.cnt0.push_back(0);
      g.upper_offset.push_back(0);
      g.links0.resize((i + 1) * maxM0);
      g.vecs.resize((i + 1) * dim);
      uint32_t word;
      std::memcpy(&word, c.data(), kU32);
      std::memcpy(&g.links0[i * maxM0], c.data() + kU32, maxM0 * kU32);
      std::memcpy(&g.vecs[i * dim], c.data() + size_links_level0, vector_size);
      std::memcpy(&g.labels[i], c.data() + size_links_level0 + vector_size, kLabelBytes);
      const uint32_t count = word & kCountMask;
      g.cnt0[i] = count;
      g.deleted[i] = (word & kDeleteMark) ? 1 : 0;
      hard(count <= maxM0, "level-0 neighbor count exceeds 2*M");
      for (uint32_t j = 0; j < count; j++) {
        const uint32_t e = g.links0[i * maxM0 + j];
        hard(e < cur, "level-0 neighbor id out of range");
        soft(e != i, "level-0 self-loop");
      }
    }

    // label lookup + upper lists (hnswalg.h:1028-1099)
    std::unordered_map<uint64_t, uint64_t> label_lookup;
    label_lookup.reserve(g.labels.size());
    uint64_t blocks = 0;
    for (uint64_t i = 0; i < cur; i++) {
      auto it = label_lookup.find(g.labels[i]);
      if (it == label_lookup.end()) {
        label_lookup[g.labels[i]] = i;
      } else {
        out.duplicate_labels++;
        if (!g.deleted[i]) {
          soft(g.deleted[it->second] != 0, "duplicate live label in index");
          it->second = i;
        }
      }
      auto size_chunk = input.LoadChunk();
      if (!size_chunk.ok()) return size_chunk.status();
      hard((*size_chunk)->size() == sizeof(uint64_t), "link-list size chunk has the wrong size");
      uint64_t link_list_size;
      std::memcpy(&link_list_size, (*size_chunk)->data(), sizeof(uint64_t));
      g.upper_offset[i] = blocks;
      if (link_list_size == 0) continue;
      hard(link_list_size % stride == 0, "upper-level link-list size is not a multiple of the stride");
      const uint64_t level = link_list_size / stride;
      // the reference holds the level in an `int` (hnswalg.h:1066): a size with garbage in its upper bytes wraps
      // here and is caught by the chunk-size check below instead — same verdict, and the same words
      hard((int)level <= header.max_level, "element level exceeds max_level");
      auto link_list_chunk = input.LoadChunk();
      if (!link_list_chunk.ok()) return link_list_chunk.status();
      const std::string &c = **link_list_chunk;
      hard(c.size() == link_list_size, "upper-level link-list chunk has the wrong size");
      g.levels[i] = (int32_t)level;
      g.upper_cnt.resize(blocks + level);
      g.upper_links.resize((blocks + level) * maxM);
      for (uint64_t l = 0; l < level; l++) {
        uint32_t word;
        std::memcpy(&word, c.data() + l * stride, kU32);
        std::memcpy(&g.upper_links[(blocks + l) * maxM], c.data() + l * stride + kU32, maxM * kU32);
        g.upper_cnt[blocks + l] = word & kCountMask;
        hard((word & kCountMask) <= maxM, "upper-level neighbor count exceeds M");
      }
      blocks += level;
    }

    // global pass (hnswalg.h:1101-1128)
    if (cur > 0)
      hard(g.enterpoint < cur && g.levels[g.enterpoint] == g.max_level, "enterpoint node is not at max_level");
    for (uint64_t i = 0; i < cur; i++) {
      for (int32_t level = 1; level <= g.levels[i]; level++) {
        const uint64_t b = g.upper_offset[i] + level - 1;
        for (uint32_t j = 0; j < g.upper_cnt[b]; j++) {
          const uint32_t e = g.upper_links[b * maxM + j];
          hard(e < cur, "upper-level neighbor id out of range");
          soft(e != i, "upper-level self-loop");
          hard(g.levels[e] >= level, "upper-level neighbor is absent at that level");
        }
      }
    }
    return out;
  } catch (const LoadFailure &f) {
    return vks::InternalError(f.msg);
  } catch (const std::exception &e) {  // vector_hnsw.cc:166-170: any exception fails the load, never the process
    return vks::InternalError(std::string("HNSWLib error while loading an index: ") + e.what());
  }
}

}  // namespace valkey_search::indexes
