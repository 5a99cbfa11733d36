// io.cpp -- on-disk graph layout of the crate (src/serialize.rs:33-209):
//   <dir>/meta               JSON HNSWMeta {layer_count, build_parameters}   (serialize.rs:27-31, 52-58)
//   <dir>/comparator         user-defined in the crate (Serializable, lib.rs:76-83); here a header
//                            {tag, metric, dim, count} followed by the raw f32 rows
//   <dir>/layer.meta.N       JSON LayerMeta {node_count, neighborhood_size}  (serialize.rs:21-25)
//   <dir>/layer.nodes.N      node_count raw native-endian u64 VectorIds      (serialize.rs:88-104)
//   <dir>/layer.neighbors.N  node_count * neighborhood_size raw u64 NodeIds  (serialize.rs:106-121)
// N counts from the bottom (serialize.rs:67) while layers[0] is the top.  The JSON is emitted
// field for field in serde's declaration order, f32 in shortest round-trip form, so a `meta`
// written here is byte-identical to the crate's for the same parameters.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <exception>
#include <string>
#include <vector>

#include "internal.h"

namespace phnsw {

static const uint64_t kComparatorTag = 0x3142574e53485042ULL;

// shortest decimal that round-trips an f32, printed the way serde_json (ryu) does for the
// magnitudes parameters take: plain decimal with at least one fractional digit
static std::string json_f32(float v) {
  if (!isfinite(v)) return "null";  // serde_json emits null for non-finite floats
  char buf[64];
  int prec = 1;
  for (; prec <= 9; prec++) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)v);
    if (strtof(buf, nullptr) == v) break;
  }
  // buf = d[.ddd]e[+-]XX
  std::string digits;
  int exp10 = 0;
  bool neg = false;
  {
    const char *p = buf;
    if (*p == '-') { neg = true; p++; }
    for (; *p && *p != 'e'; p++)
      if (*p != '.') digits.push_back(*p);
    if (*p == 'e') exp10 = atoi(p + 1);
  }
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string out = neg ? "-" : "";
  if (v == 0.0f) return out + "0.0";
  int nd = (int)digits.size();
  if (exp10 >= -5 && exp10 < 16) {
    if (exp10 < 0) {
      out += "0.";
      out.append((size_t)(-exp10 - 1), '0');
      out += digits;
    } else if (exp10 + 1 >= nd) {
      out += digits;
      out.append((size_t)(exp10 + 1 - nd), '0');
      out += ".0";
    } else {
      out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
    }
  } else {
    out += digits.substr(0, 1);
    if (nd > 1) out += "." + digits.substr(1);
    out += "e" + std::to_string(exp10);
  }
  return out;
}

static std::string json_search_params(const phnsw_search_params &sp) {
  char b[256];
  snprintf(b, sizeof b,
           "{\"number_of_candidates\":%llu,\"upper_layer_candidate_count\":%llu,\"probe_depth\":%llu}",
           (unsigned long long)sp.number_of_candidates,
           (unsigned long long)sp.upper_layer_candidate_count, (unsigned long long)sp.probe_depth);
  return b;
}

std::string json_build_params(const phnsw_build_params &bp) {
  std::string s = "{\"order\":" + std::to_string(bp.order) +
                  ",\"zero_layer_neighborhood_size\":" + std::to_string(bp.zero_layer_neighborhood_size) +
                  ",\"neighborhood_size\":" + std::to_string(bp.neighborhood_size) +
                  ",\"optimization\":{\"promotion_threshold\":" + json_f32(bp.optimization.promotion_threshold) +
                  ",\"neighborhood_threshold\":" + json_f32(bp.optimization.neighborhood_threshold) +
                  ",\"recall_proportion\":" + json_f32(bp.optimization.recall_proportion) +
                  ",\"promotion_proportion\":" + json_f32(bp.optimization.promotion_proportion) +
                  ",\"search\":" + json_search_params(bp.optimization.search) +
                  "},\"initial_partition_search\":" + json_search_params(bp.initial_partition_search) + "}";
  return s;
}

static bool write_all(const std::string &path, const void *buf, size_t len) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  size_t w = len ? fwrite(buf, 1, len, f) : 0;
  bool ok = (w == len);
  if (fclose(f) != 0) ok = false;
  return ok;
}
static bool read_all(const std::string &path, std::vector<char> &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  long sz = -1;
  if (fseek(f, 0, SEEK_END) == 0) sz = ftell(f);
  // the documents read this way are small JSON files: anything else is not one of ours
  if (sz < 0 || sz > (64L << 20) || fseek(f, 0, SEEK_SET) != 0) {
    fclose(f);
    errno = sz < 0 ? errno : EFBIG;
    return false;
  }
  out.assign((size_t)sz + 1, 0);
  size_t r = sz ? fread(out.data(), 1, (size_t)sz, f) : 0;
  fclose(f);
  return r == (size_t)sz;
}

// ---- a small JSON reader for the flat numeric objects serde_json emits ----
struct JsonCur {
  const char *p;
  bool ok = true;
  void ws() { while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') p++; }
  bool eat(char c) { ws(); if (*p == c) { p++; return true; } return false; }
  std::string str() {
    ws();
    std::string s;
    if (*p != '"') { ok = false; return s; }
    p++;
    while (*p && *p != '"') { if (*p == '\\' && p[1]) p++; s.push_back(*p++); }
    if (*p == '"') p++; else ok = false;
    return s;
  }
  double num() {
    ws();
    char *e;
    double v = strtod(p, &e);
    if (e == p) ok = false;
    p = e;
    return v;
  }
  void skip() {  // skip any value
    ws();
    if (*p == '{') { p++; if (eat('}')) return; do { str(); eat(':'); skip(); } while (ok && eat(',')); if (!eat('}')) ok = false; }
    else if (*p == '[') { p++; if (eat(']')) return; do { skip(); } while (ok && eat(',')); if (!eat(']')) ok = false; }
    else if (*p == '"') str();
    else if (!strncmp(p, "true", 4)) p += 4;
    else if (!strncmp(p, "false", 5)) p += 5;
    else if (!strncmp(p, "null", 4)) p += 4;
    else num();
  }
};

static bool parse_search_params(JsonCur &c, phnsw_search_params *sp) {
  int seen = 0;
  if (!c.eat('{')) return false;
  if (!c.eat('}')) {
    do {
      std::string k = c.str();
      if (!c.eat(':')) return false;
      if (k == "number_of_candidates") { sp->number_of_candidates = (uint64_t)c.num(); seen |= 1; }
      else if (k == "upper_layer_candidate_count") { sp->upper_layer_candidate_count = (uint64_t)c.num(); seen |= 2; }
      else if (k == "probe_depth") { sp->probe_depth = (uint64_t)c.num(); seen |= 4; }
      else c.skip();
    } while (c.ok && c.eat(','));
    if (!c.eat('}')) return false;
  }
  return c.ok && seen == 7;  // serde: missing field is an error
}
static bool parse_optimization(JsonCur &c, phnsw_optimization_params *op) {
  int seen = 0;
  if (!c.eat('{')) return false;
  if (!c.eat('}')) {
    do {
      std::string k = c.str();
      if (!c.eat(':')) return false;
      if (k == "promotion_threshold") { op->promotion_threshold = (float)c.num(); seen |= 1; }
      else if (k == "neighborhood_threshold") { op->neighborhood_threshold = (float)c.num(); seen |= 2; }
      else if (k == "recall_proportion") { op->recall_proportion = (float)c.num(); seen |= 4; }
      else if (k == "promotion_proportion") { op->promotion_proportion = (float)c.num(); seen |= 8; }
      else if (k == "search") { if (!parse_search_params(c, &op->search)) return false; seen |= 16; }
      else c.skip();
    } while (c.ok && c.eat(','));
    if (!c.eat('}')) return false;
  }
  return c.ok && seen == 31;
}
bool parse_build_params(JsonCur &c, phnsw_build_params *bp) {
  int seen = 0;
  if (!c.eat('{')) return false;
  if (!c.eat('}')) {
    do {
      std::string k = c.str();
      if (!c.eat(':')) return false;
      if (k == "order") { bp->order = (uint64_t)c.num(); seen |= 1; }
      else if (k == "zero_layer_neighborhood_size") { bp->zero_layer_neighborhood_size = (uint64_t)c.num(); seen |= 2; }
      else if (k == "neighborhood_size") { bp->neighborhood_size = (uint64_t)c.num(); seen |= 4; }
      else if (k == "optimization") { if (!parse_optimization(c, &bp->optimization)) return false; seen |= 8; }
      else if (k == "initial_partition_search") { if (!parse_search_params(c, &bp->initial_partition_search)) return false; seen |= 16; }
      else c.skip();
    } while (c.ok && c.eat(','));
    if (!c.eat('}')) return false;
  }
  return c.ok && seen == 31;
}
static bool parse_meta(const char *s, uint64_t *layer_count, phnsw_build_params *bp) {
  JsonCur c{s};
  int seen = 0;
  if (!c.eat('{')) return false;
  if (!c.eat('}')) {
    do {
      std::string k = c.str();
      if (!c.eat(':')) return false;
      if (k == "layer_count") { *layer_count = (uint64_t)c.num(); seen |= 1; }
      else if (k == "build_parameters") { if (!parse_build_params(c, bp)) return false; seen |= 2; }
      else c.skip();
    } while (c.ok && c.eat(','));
    if (!c.eat('}')) return false;
  }
  return c.ok && seen == 3;
}
static bool parse_layer_meta(const char *s, uint64_t *node_count, uint64_t *M) {
  JsonCur c{s};
  int seen = 0;
  if (!c.eat('{')) return false;
  if (!c.eat('}')) {
    do {
      std::string k = c.str();
      if (!c.eat(':')) return false;
      if (k == "node_count") { *node_count = (uint64_t)c.num(); seen |= 1; }
      else if (k == "neighborhood_size") { *M = (uint64_t)c.num(); seen |= 2; }
      else c.skip();
    } while (c.ok && c.eat(','));
    if (!c.eat('}')) return false;
  }
  return c.ok && seen == 3;
}

static int mkdir_p(const std::string &dir) {  // create_dir_all (serialize.rs:40)
  std::string cur;
  for (size_t i = 0; i <= dir.size(); i++) {
    if (i == dir.size() || dir[i] == '/') {
      if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return -1;
    }
    if (i < dir.size()) cur.push_back(dir[i]);
  }
  return 0;
}


int io_mkdir_p(const std::string &dir) { return mkdir_p(dir); }

// <dir>/meta + <dir>/layer.{meta,nodes,neighbors}.N (serialize.rs:33-124 minus the comparator)
phnsw_status io_save_graph(const phnsw_index *ix, const std::string &d) {
  if (mkdir_p(d) != 0) {
    set_error("save: cannot create %s: %s", d.c_str(), strerror(errno));
    return PHNSW_ERR_IO;
  }
  const uint64_t L = ix->layers.size();
  std::string meta = "{\"layer_count\":" + std::to_string(L) +
                     ",\"build_parameters\":" + json_build_params(ix->bp) + "}";
  if (!write_all(d + "/meta", meta.data(), meta.size())) {
    set_error("save: cannot write %s/meta: %s", d.c_str(), strerror(errno));
    return PHNSW_ERR_IO;
  }
  for (uint64_t i = 0; i < L; i++) {
    const uint64_t num = L - i - 1;
    const LayerStore &l = ix->layers[i];
    std::string lm = "{\"node_count\":" + std::to_string(l.node_count) +
                     ",\"neighborhood_size\":" + std::to_string(l.M) + "}";
    std::vector<uint64_t> nodes(l.node_count), nb((size_t)l.node_count * l.M);
    phnsw_status rc = phnsw_index_export_layer(ix, i, nodes.data(), nb.data());
    if (rc != PHNSW_OK) return rc;
    std::string n = std::to_string(num);
    if (!write_all(d + "/layer.meta." + n, lm.data(), lm.size()) ||
        !write_all(d + "/layer.nodes." + n, nodes.data(), nodes.size() * 8) ||
        !write_all(d + "/layer.neighbors." + n, nb.data(), nb.size() * 8)) {
      set_error("save: cannot write layer %llu under %s: %s", (unsigned long long)num, d.c_str(),
                strerror(errno));
      return PHNSW_ERR_IO;
    }
  }
  return PHNSW_OK;
}

// the comparator file: {tag, metric, dim, count} + raw f32 rows
phnsw_status io_save_store(const phnsw_store *s, const std::string &path) {
  std::vector<float> rows((size_t)s->n * s->dim);
  std::vector<uint64_t> ids(s->n);
  for (uint64_t i = 0; i < s->n; i++) ids[i] = i;
  phnsw_status rc = phnsw_store_get_rows(s, ids.data(), s->n, rows.data());
  if (rc != PHNSW_OK) return rc;
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    set_error("save: cannot write %s: %s", path.c_str(), strerror(errno));
    return PHNSW_ERR_IO;
  }
  uint64_t hdr[4] = {kComparatorTag, (uint64_t)s->metric, s->dim, s->n};
  bool ok = fwrite(hdr, sizeof hdr, 1, f) == 1;
  if (ok && !rows.empty()) ok = fwrite(rows.data(), 4, rows.size(), f) == rows.size();
  if (fclose(f) != 0) ok = false;
  if (!ok) {
    set_error("save: short write on %s", path.c_str());
    return PHNSW_ERR_IO;
  }
  return PHNSW_OK;
}

phnsw_status io_load_store(const std::string &path, int device, phnsw_store **out) {
  *out = nullptr;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {  // serialize.rs:143-145
    set_error("Index not found");
    return PHNSW_ERR_NOT_FOUND;
  }
  uint64_t hdr[4];
  if (fread(hdr, sizeof hdr, 1, f) != 1 || hdr[0] != kComparatorTag || hdr[1] > 3 || hdr[2] == 0) {
    fclose(f);
    set_error("load: %s has an unknown header", path.c_str());
    return PHNSW_ERR_FORMAT;
  }
  // dim * count comes from the file: check it against the file's own size before allocating
  struct stat sb;
  const bool sized = fstat(fileno(f), &sb) == 0;
  const unsigned __int128 need = (unsigned __int128)hdr[2] * hdr[3] * 4 + sizeof hdr;
  if (!sized || hdr[2] > (1ull << 20) || need > (unsigned __int128)(uint64_t)sb.st_size) {
    fclose(f);
    set_error("load: %s is truncated (header names %llu x %llu floats)", path.c_str(),
              (unsigned long long)hdr[3], (unsigned long long)hdr[2]);
    return PHNSW_ERR_IO;
  }
  std::vector<float> rows;
  try {
    rows.resize((size_t)hdr[2] * hdr[3]);
  } catch (const std::exception &) {
    fclose(f);
    set_error("load: out of host memory for %s", path.c_str());
    return PHNSW_ERR_IO;
  }
  size_t got = rows.empty() ? 0 : fread(rows.data(), 4, rows.size(), f);
  fclose(f);
  if (got != rows.size()) {
    set_error("load: %s is truncated", path.c_str());
    return PHNSW_ERR_IO;
  }
  return phnsw_store_create((phnsw_metric)hdr[1], hdr[2], hdr[3], rows.data(), device, out);
}

// layers of <dir> over an existing store
phnsw_status io_load_graph(const std::string &d, phnsw_store *s, phnsw_index **out) {
  *out = nullptr;
  const char *dir = d.c_str();
  std::vector<char> buf;
  if (!read_all(d + "/meta", buf)) {
    set_error("load: cannot read %s/meta: %s", dir, strerror(errno));
    return PHNSW_ERR_IO;
  }
  uint64_t L = 0;
  phnsw_build_params bp;
  phnsw_default_build_params(&bp);
  if (!parse_meta(buf.data(), &L, &bp)) {
    set_error("load: %s/meta is not a valid HNSWMeta document", dir);
    return PHNSW_ERR_FORMAT;
  }
  // layer.nodes.N / layer.neighbors.N are raw little-endian u64 arrays (serialize.rs:88-121):
  // they are mapped read-only and handed to the uploader as they lie in the page cache, which
  // compacts them to u32 on the device -- no intermediate host copy
  struct Mapping {
    void *p = nullptr;
    size_t len = 0;
    ~Mapping() { if (p) munmap(p, len); }
  };
  // sizes come from the files: bound them before anything is allocated from them (the crate
  // builds at most a few tens of layers: node counts shrink by `order` per layer)
  if (L > 64) {
    set_error("load: %s/meta names %llu layers (at most 64 supported)", dir, (unsigned long long)L);
    return PHNSW_ERR_FORMAT;
  }
  std::vector<Mapping> maps(2 * L);
  std::vector<phnsw_layer_desc> descs(L);
  for (uint64_t i = 0; i < L; i++) {
    std::string n = std::to_string(L - i - 1);
    uint64_t nc = 0, M = 0;
    if (!read_all(d + "/layer.meta." + n, buf)) {
      set_error("load: cannot read %s/layer.meta.%s: %s", dir, n.c_str(), strerror(errno));
      return PHNSW_ERR_IO;
    }
    if (!parse_layer_meta(buf.data(), &nc, &M)) {
      set_error("load: %s/layer.meta.%s is not a valid LayerMeta document", dir, n.c_str());
      return PHNSW_ERR_FORMAT;
    }
    if (nc >= 0x7FFFFFFFull || M > 64) {  // 32-bit node ids on the device; rows of at most 64
      set_error("load: %s/layer.meta.%s: node_count %llu / neighborhood_size %llu out of range", dir,
                n.c_str(), (unsigned long long)nc, (unsigned long long)M);
      return PHNSW_ERR_FORMAT;
    }
    const uint64_t *ptr[2] = {nullptr, nullptr};
    for (int which = 0; which < 2; which++) {
      const size_t want = (size_t)(which ? nc * M : nc) * 8;  // < 2^31 * 64 * 8: no overflow
      std::string p = d + (which ? "/layer.neighbors." : "/layer.nodes.") + n;
      int fd = open(p.c_str(), O_RDONLY);
      struct stat sb;
      bool ok = fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size >= want;  // read_exact
      if (ok && want) {
        Mapping &m = maps[2 * i + which];
        void *q = mmap(nullptr, want, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (q == MAP_FAILED) ok = false;
        else {
          m.p = q;
          m.len = want;
          ptr[which] = (const uint64_t *)q;
        }
      }
      if (fd >= 0) close(fd);
      if (!ok) {  // a missing or short file is an io error
        set_error("load: cannot read %s", p.c_str());
        return PHNSW_ERR_IO;
      }
    }
    descs[i].node_count = nc;
    descs[i].neighborhood_size = M;
    descs[i].nodes = ptr[0];
    descs[i].neighbors = ptr[1];
  }
  return phnsw_index_from_layers(s, L, descs.data(), &bp, out);
}

// pq_build_parameters.json (pq.rs:94-117): serde_json of PqBuildParameters (parameters.rs:66-71)
phnsw_status io_save_pq_params(const std::string &path, const phnsw_pq_build_params &bp) {
  std::string j = "{\"centroids\":" + json_build_params(bp.centroids) + ",\"hnsw\":" +
                  json_build_params(bp.hnsw) + ",\"quantized_search\":" +
                  json_search_params(bp.quantized_search) + "}";
  if (!write_all(path, j.data(), j.size())) {
    set_error("save: cannot write %s: %s", path.c_str(), strerror(errno));
    return PHNSW_ERR_IO;
  }
  return PHNSW_OK;
}
phnsw_status io_load_pq_params(const std::string &path, phnsw_pq_build_params *bp) {
  std::vector<char> buf;
  if (!read_all(path, buf)) {
    set_error("load: cannot read %s: %s", path.c_str(), strerror(errno));
    return PHNSW_ERR_IO;
  }
  JsonCur c{buf.data()};
  int seen = 0;
  bool ok = c.eat('{');
  if (ok && !c.eat('}')) {
    do {
      std::string k = c.str();
      if (!c.eat(':')) { ok = false; break; }
      if (k == "centroids") { ok = parse_build_params(c, &bp->centroids); seen |= 1; }
      else if (k == "hnsw") { ok = parse_build_params(c, &bp->hnsw); seen |= 2; }
      else if (k == "quantized_search") { ok = parse_search_params(c, &bp->quantized_search); seen |= 4; }
      else c.skip();
    } while (ok && c.ok && c.eat(','));
    if (ok && !c.eat('}')) ok = false;
  }
  if (!ok || !c.ok || seen != 7) {
    set_error("load: %s is not a valid PqBuildParameters document", path.c_str());
    return PHNSW_ERR_FORMAT;
  }
  return PHNSW_OK;
}

}  // namespace phnsw

using namespace phnsw;

extern "C" {

phnsw_status phnsw_format_build_params(const phnsw_build_params *bp, char *out, uint64_t out_cap) {
  if (!bp || !out) return PHNSW_ERR_INVALID;
  std::string s = json_build_params(*bp);
  if (s.size() + 1 > out_cap) return PHNSW_ERR_INVALID;
  memcpy(out, s.c_str(), s.size() + 1);
  return PHNSW_OK;
}

phnsw_status phnsw_index_save(const phnsw_index *ix, const char *dir) {
  PH_ENTRY();
  if (!ix || !dir) return PHNSW_ERR_INVALID;
  if (!ix->store->rows) {
    set_error("save: not available on a PQ8 store");
    return PHNSW_ERR_INVALID;
  }
  std::string d(dir);
  phnsw_status rc = io_save_graph(ix, d);
  if (rc != PHNSW_OK) return rc;
  // the comparator entry is only written for a non-empty index (serialize.rs:60-65)
  if (!ix->layers.empty()) rc = io_save_store(ix->store, d + "/comparator");
  return rc;
}

phnsw_status phnsw_index_load(const char *dir, int device, phnsw_store **store_out,
                              phnsw_index **index_out) {
  PH_ENTRY();
  if (!dir || !store_out || !index_out) return PHNSW_ERR_INVALID;
  *store_out = nullptr;
  *index_out = nullptr;
  std::string d(dir);
  std::vector<char> buf;
  if (!read_all(d + "/meta", buf)) {  // read before the comparator, as serialize.rs:130-145 does
    set_error("load: cannot read %s/meta: %s", dir, strerror(errno));
    return PHNSW_ERR_IO;
  }
  phnsw_store *s = nullptr;
  phnsw_status rc;
  phnsw_index *ix = nullptr;
  try {  // nothing may unwind across the C ABI
    rc = io_load_store(d + "/comparator", device, &s);
    if (rc != PHNSW_OK) return rc;
    rc = io_load_graph(d, s, &ix);
  } catch (const std::exception &e) {
    set_error("load: %s", e.what());
    rc = PHNSW_ERR_IO;
  }
  if (rc != PHNSW_OK) {
    phnsw_store_destroy(s);
    return rc;
  }
  *store_out = s;
  *index_out = ix;
  return PHNSW_OK;
}

}  // extern "C"
