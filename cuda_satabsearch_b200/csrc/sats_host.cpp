// sats_host.cpp -- host half of libsats: structure parsing, database container, file formats,
// Gumbel statistics, result formatting, sharding.  No CUDA here.
//
// Behaviour follows the reference's host libraries (stivalaa/cuda_satabsearch, nvcc_src_current/):
//   parsetableaux.c:52-138   SSE-type and tableau-code encodings
//   parsetableaux.c:193-294  row grammar of tableau / distance-matrix blocks (3- and 7-column cells)
//   parsetableaux.c:317-632  read_database / read_queries (header "%8s %d", oversize entries skipped)
//   gumbelstats.c:50-94      norm2 / z_gumbel / pv_gumbel
//   cudaSaTabsearch.cu:415-454, 631-693  output grammar, stdin grammar
//   scripts/convdb2.py:182-231  ASCII database writer
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <numeric>
#include <thread>
#include <strings.h>

#include "sats_internal.h"

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

int sats_fail(int status, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return status;
}

extern "C" const char *sats_last_error(void) { return g_err; }
extern "C" const char *sats_version(void) { return "cuda_satabsearch_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------ container
void sats_db::append(const char *nm, int n, const uint8_t *tri_tab, const float *tri_dmat)
{
  if (tri_off.empty()) tri_off.push_back(0);
  int64_t cells = (int64_t)n * (n + 1) / 2;
  order.push_back(n);
  char slot[9] = {0};
  strncpy(slot, nm, SATS_LABELSIZE);
  names.insert(names.end(), slot, slot + 9);
  tab.insert(tab.end(), tri_tab, tri_tab + cells);
  dmat.insert(dmat.end(), tri_dmat, tri_dmat + cells);
  tri_off.push_back(tri_off.back() + cells);
}

// ------------------------------------------------------------------------------------------ parsing
namespace {

struct Cursor {
  const char *p, *end;
  bool eof() const { return p >= end; }
  // next line without its terminator; false at end of text
  bool line(const char *&b, const char *&e)
  {
    if (p >= end) return false;
    b = p;
    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
    e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    return true;
  }
  void skip_space()
  {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) p++;
  }
};

int type_code(char c0, char c1, int *out)
{
  if (c0 == 'e') { *out = 0; return 0; }
  switch (c1) {
    case 'a': *out = 1; return 0;
    case 'i': *out = 2; return 0;
    case 'g': *out = 3; return 0;
    default: return sats_fail(SATS_ERR_PARSE, "Bad helix type %c", c1);
  }
}

int tableau_code(char c0, char c1, int *out)
{
  int hi, lo;
  switch (c0) {
    case 'P': hi = 0; break;
    case 'R': hi = 1; break;
    case 'O': hi = 2; break;
    case 'L': hi = 3; break;
    case '?': hi = 4; break;
    default: return sats_fail(SATS_ERR_PARSE, "invalid tableaux code %c", c0);
  }
  switch (c1) {
    case 'E': lo = 0; break;
    case 'D': lo = 1; break;
    case 'S': lo = 2; break;
    case 'T': lo = 3; break;
    case '?': lo = 4; break;
    default: return sats_fail(SATS_ERR_PARSE, "invalid tableaux code %c", c1);
  }
  *out = (hi << 4) | lo;
  return 0;
}

inline bool is_digit(char ch) { return ch >= '0' && ch <= '9'; }

// strtof("dd.ddd") for all 100 000 texts the "%6.3f" format can produce below 100, indexed by the five digits
const float *canonical_distances()
{
  static const std::vector<float> table = [] {
    std::vector<float> t(100000);
    char buf[16];
    for (int k = 0; k < 100000; k++) {
      snprintf(buf, sizeof buf, "%d.%03d", k / 1000, k % 1000);
      t[(size_t)k] = strtof(buf, nullptr);
    }
    return t;
  }();
  return table.data();
}

// header token pair "%8s %d": a name of at most 8 characters, then the order
bool read_header(Cursor &c, char name[9], int *order)
{
  c.skip_space();
  if (c.eof()) return false;
  int k = 0;
  while (c.p < c.end && k < 8 && !isspace((unsigned char)*c.p)) name[k++] = *c.p++;
  name[k] = 0;
  if (k == 0) return false;
  c.skip_space();
  if (c.eof()) return false;
  char num[16];
  int m = 0;
  if (c.p < c.end && (*c.p == '-' || *c.p == '+')) num[m++] = *c.p++;
  while (c.p < c.end && m < 15 && *c.p >= '0' && *c.p <= '9') num[m++] = *c.p++;
  num[m] = 0;
  if (m == 0 || (m == 1 && (num[0] == '-' || num[0] == '+'))) return false;
  *order = atoi(num);
  // the trailing "\n" of the reference's fscanf format eats the rest of the whitespace
  while (c.p < c.end && (*c.p == ' ' || *c.p == '\t' || *c.p == '\r')) c.p++;
  if (c.p < c.end && *c.p == '\n') c.p++;
  return true;
}

// Parses consecutive entries until the text ends or a header fails to parse.
// `warn` (parallel parsing): collect the oversize warnings instead of printing them, and leave the summary line to the caller
int parse_entries(Cursor &c, const char *what, sats_db *db, int max_order = SATS_MAXDIM, std::string *warn = nullptr,
                  int *skipped_out = nullptr)
{
  std::vector<uint8_t> ttab;
  std::vector<float> tdm;
  int skipped = 0;
  char name[9];
  int n;
  while (read_header(c, name, &n)) {
    if (n < 1) return sats_fail(SATS_ERR_PARSE, "%s structure %s has bad order %d", what, name, n);
    bool keep = n <= max_order;
    if (!keep) {
      char msg[256];
      int m = snprintf(msg, sizeof msg, "Tableau %s order %d is too large (max is %d)\nWARNING: excluded %s structure %s as it is too large\n",
                       name, n, max_order, what, name);
      if (warn) warn->append(msg, (size_t)m);
      else fputs(msg, stderr);
      skipped++;
    }
    size_t cells = (size_t)n * (n + 1) / 2;
    ttab.assign(cells, 0);
    tdm.assign(cells, 0.f);
    const char *b, *e;
    for (int i = 0; i < n; i++) {
      if (!c.line(b, e)) return sats_fail(SATS_ERR_PARSE, "%s structure %s: truncated tableau (row %d of %d)", what, name, i + 1, n);
      if (!keep) continue;
      if (e - b < 3 * i + 2) return sats_fail(SATS_ERR_PARSE, "%s structure %s: tableau row %d too short", what, name, i + 1);
      for (int j = 0; j <= i; j++) {
        int v, rc;
        rc = (i == j) ? type_code(b[3 * j], b[3 * j + 1], &v) : tableau_code(b[3 * j], b[3 * j + 1], &v);
        if (rc) return rc;
        ttab[(size_t)i * (i + 1) / 2 + j] = (uint8_t)v;
      }
    }
    char field[64];
    const float *canon = canonical_distances();
    for (int i = 0; i < n; i++) {
      if (!c.line(b, e)) return sats_fail(SATS_ERR_PARSE, "%s structure %s: truncated distance matrix (row %d of %d)", what, name, i + 1, n);
      if (!keep) continue;
      for (int j = 0; j <= i; j++) {
        // strtof(&buf[j*7]): conversion starts at column 7j and stops at the first character that
        // cannot continue a number (so a wider value bleeds into the next field, as in the reference)
        const char *s = b + 7 * j;
        if (s >= e) return sats_fail(SATS_ERR_PARSE, "%s structure %s: distance row %d too short", what, name, i + 1);
        // the writer's own "%6.3f " shape -- "[ d]d.ddd" and then something that ends the number: what strtof returns for
        // it was tabulated once (with strtof), so this is the same float without the call
        if (e - s >= 6 && s[2] == '.' && (s[0] == ' ' || is_digit(s[0])) && is_digit(s[1]) && is_digit(s[3]) && is_digit(s[4]) &&
            is_digit(s[5]) && (e - s == 6 || s[6] == ' ' || s[6] == '\t' || s[6] == '\r')) {
          const int k = (s[0] == ' ' ? 0 : (s[0] - '0') * 10000) + (s[1] - '0') * 1000 + (s[3] - '0') * 100 + (s[4] - '0') * 10 + (s[5] - '0');
          tdm[(size_t)i * (i + 1) / 2 + j] = canon[k];
          continue;
        }
        size_t len = std::min((size_t)(e - s), sizeof field - 1);
        memcpy(field, s, len);
        field[len] = 0;
        tdm[(size_t)i * (i + 1) / 2 + j] = strtof(field, nullptr);
      }
    }
    if (keep) db->append(name, n, ttab.data(), tdm.data());
  }
  if (skipped_out) *skipped_out = skipped;
  else if (skipped) fprintf(stderr, "WARNING: skipped %d %s tableaux of order > %d\n", skipped, what, max_order);
  return 0;
}

// A large database text parsed on several threads: entries are independent, and in the writer's format a blank line separates
// them, so the text is cut at blank lines near the even split points and every piece goes through parse_entries() on its own.
// Returns false -- and the caller parses serially, which also reproduces the serial error message -- whenever anything is
// not plainly regular: no blank-line boundary near a split point, a piece that fails or stops before its end.
bool parse_entries_parallel(const char *text, size_t len, int max_order, sats_db *db)
{
  size_t want = 8;
  if (const char *e = getenv("SATS_PARSE_THREADS")) want = (size_t)std::max(1, atoi(e));
  const size_t threads = std::min<size_t>(want, std::max(1u, std::thread::hardware_concurrency()));
  if (threads < 2 || len < (size_t)(4u << 20)) return false;
  std::vector<size_t> cut(1, 0);
  for (size_t t = 1; t < threads; t++) {
    size_t pos = std::max(cut.back(), len / threads * t);
    const char *hit = nullptr;
    for (const char *q = text + pos; q + 1 < text + len; q++)
      if (q[0] == '\n' && q[1] == '\n') { hit = q + 2; break; }
    if (!hit) break;
    // the piece must start like a header: a name, an integer, end of line
    const char *q = hit;
    while (q < text + len && *q != '\n' && !isspace((unsigned char)*q)) q++;
    if (q == hit) return false;
    while (q < text + len && (*q == ' ' || *q == '\t')) q++;
    const char *num = q;
    while (q < text + len && *q >= '0' && *q <= '9') q++;
    if (q == num) return false;
    while (q < text + len && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
    if (q < text + len && *q != '\n') return false;
    if ((size_t)(hit - text) > cut.back()) cut.push_back((size_t)(hit - text));
  }
  cut.push_back(len);
  const size_t pieces = cut.size() - 1;
  if (pieces < 2) return false;
  struct Piece { sats_db db; std::string warn; int skipped = 0, rc = 0; bool complete = false; };
  std::vector<Piece> out(pieces);
  {
    std::vector<std::thread> pool;
    struct Join { std::vector<std::thread> &p; ~Join() { for (auto &t : p) if (t.joinable()) t.join(); } } join{pool};
    auto work = [&](size_t k) {
      Piece &pc = out[k];
      try {
        pc.db.tri_off.push_back(0);
        Cursor c{text + cut[k], text + cut[k + 1]};
        pc.rc = parse_entries(c, "database", &pc.db, max_order, &pc.warn, &pc.skipped);
        c.skip_space();
        pc.complete = pc.rc == 0 && c.eof();
      } catch (...) { pc.rc = SATS_ERR_NOMEM; }
    };
    for (size_t k = 1; k < pieces; k++) pool.emplace_back(work, k);
    work(0);
  }
  for (const Piece &pc : out) if (!pc.complete) return false;
  int skipped = 0;
  size_t n_entries = 0, n_cells = 0;
  for (const Piece &pc : out) { n_entries += pc.db.order.size(); n_cells += pc.db.tab.size(); }
  db->order.reserve(n_entries); db->names.reserve(9 * n_entries); db->tri_off.reserve(n_entries + 1);
  db->tab.reserve(n_cells); db->dmat.reserve(n_cells);
  for (Piece &pc : out) {
    fputs(pc.warn.c_str(), stderr);
    skipped += pc.skipped;
    const int64_t base = db->tri_off.back();
    db->order.insert(db->order.end(), pc.db.order.begin(), pc.db.order.end());
    db->names.insert(db->names.end(), pc.db.names.begin(), pc.db.names.end());
    db->tab.insert(db->tab.end(), pc.db.tab.begin(), pc.db.tab.end());
    db->dmat.insert(db->dmat.end(), pc.db.dmat.begin(), pc.db.dmat.end());
    for (size_t e = 1; e < pc.db.tri_off.size(); e++) db->tri_off.push_back(base + pc.db.tri_off[e]);
  }
  if (skipped) fprintf(stderr, "WARNING: skipped %d %s tableaux of order > %d\n", skipped, "database", max_order);
  return true;
}

// a structure handed over as arrays or read from the packed cache must use the same alphabet the ASCII parser enforces:
// SSE type 0..3 on the diagonal, letters 0..4 in both nibbles elsewhere (the kernel indexes tables with them)
int check_codes(const char *what, const char *name, int n, const uint8_t *tri_tab)
{
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      const uint8_t v = tri_tab[(size_t)i * (i + 1) / 2 + j];
      if (i == j ? v > 3 : ((v >> 4) > 4 || (v & 15) > 4))
        return sats_fail(SATS_ERR_PARSE, "%s: structure %.8s has invalid %s 0x%02x at (%d, %d)", what, name,
                         i == j ? "SSE type" : "tableau code", v, i, j);
    }
  return 0;
}

int slurp(const char *path, std::string *out)
{
  FILE *fp = fopen(path, "rb");
  if (!fp) return sats_fail(SATS_ERR_IO, "ERROR opening db file %s", path);
  // regular file: one allocation, one read; anything else (pipe, /proc): chunked
  if (fseek(fp, 0, SEEK_END) == 0) {
    const long size = ftell(fp);
    if (size > 0 && fseek(fp, 0, SEEK_SET) == 0) {
      out->resize((size_t)size);
      const size_t got = fread(&(*out)[0], 1, (size_t)size, fp);
      out->resize(got);
    } else {
      rewind(fp);
    }
  }
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, fp)) > 0) out->append(buf, n);
  fclose(fp);
  return 0;
}

}  // namespace

extern "C" int sats_db_parse_ascii_ext(const char *text, size_t len, int max_order, sats_db **out)
try {
  if (!text || !out) return sats_fail(SATS_ERR_ARG, "sats_db_parse_ascii: null argument");
  if (max_order < 1 || max_order > SATS_MAXDIM_EXT) return sats_fail(SATS_ERR_ARG, "max_order %d outside 1..%d", max_order, SATS_MAXDIM_EXT);
  std::unique_ptr<sats_db> db(new sats_db());
  db->tri_off.push_back(0);
  if (!parse_entries_parallel(text, len, max_order, db.get())) {
    db.reset(new sats_db());
    db->tri_off.push_back(0);
    Cursor c{text, text + len};
    int rc = parse_entries(c, "database", db.get(), max_order);
    if (rc) return rc;
  }
  *out = db.release();
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" int sats_db_parse_ascii(const char *text, size_t len, sats_db **out)
{
  return sats_db_parse_ascii_ext(text, len, SATS_MAXDIM, out);
}

extern "C" int sats_db_read_ascii_ext(const char *path, int max_order, sats_db **out)
try {
  if (!path || !out) return sats_fail(SATS_ERR_ARG, "sats_db_read_ascii: null argument");
  std::string text;
  int rc = slurp(path, &text);
  if (rc) return rc;
  return sats_db_parse_ascii_ext(text.data(), text.size(), max_order, out);
}
SATS_CATCH_ALL

extern "C" int sats_db_read_ascii(const char *path, sats_db **out) { return sats_db_read_ascii_ext(path, SATS_MAXDIM, out); }

extern "C" int sats_input_parse(const char *text, size_t len, char *dbfile, size_t dbfile_cap, int flags_tf[3],
                                sats_db **queries)
try {
  if (!text || !dbfile || !flags_tf || !queries || dbfile_cap < 2) return sats_fail(SATS_ERR_ARG, "sats_input_parse: bad argument");
  Cursor c{text, text + len};
  c.skip_space();
  size_t k = 0;
  while (c.p < c.end && !isspace((unsigned char)*c.p)) {
    if (k + 1 < dbfile_cap) dbfile[k++] = *c.p;
    c.p++;
  }
  dbfile[k] = 0;
  if (k == 0) return sats_fail(SATS_ERR_PARSE, "ERROR reading dbfilename from stdin");
  c.skip_space();
  // "%c %c %c\n"
  char f[3];
  for (int i = 0; i < 3; i++) {
    if (i) c.skip_space();
    if (c.eof()) return sats_fail(SATS_ERR_PARSE, "ERROR reading options from stdin");
    f[i] = *c.p++;
  }
  for (int i = 0; i < 3; i++) flags_tf[i] = (f[i] == 'T');
  std::unique_ptr<sats_db> q(new sats_db());
  q->tri_off.push_back(0);
  int rc = parse_entries(c, "query", q.get());
  if (rc) return rc;
  if (q->count() == 0) return sats_fail(SATS_ERR_PARSE, "ERROR: no query structures found on stdin");
  *queries = q.release();
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" int sats_idlist_parse(const char *text, size_t len, char *ids_out, int max_ids)
{
  if (!text || !ids_out) return sats_fail(SATS_ERR_ARG, "sats_idlist_parse: null argument");
  Cursor c{text, text + len};
  const char *b, *e;
  int n = 0;
  while (c.line(b, e)) {
    while (e > b && (e[-1] == '\r')) e--;
    if (e == b) continue;               // blank line (the reference would fail on it later)
    if (n >= max_ids) return sats_fail(SATS_ERR_ARG, "too many query ids (max %d)", max_ids);
    char *slot = ids_out + (size_t)n * 9;
    memset(slot, 0, 9);
    size_t k = std::min((size_t)(e - b), (size_t)(SATS_LABELSIZE - 1));   // queryptr[LABELSIZE-1] = '\0'
    memcpy(slot, b, k);
    n++;
  }
  return n;
}

extern "C" int sats_db_from_arrays(int count, const int32_t *order, const char *names, const int64_t *off,
                                   const uint8_t *tabs, const float *dmats, sats_db **out)
try {
  if (count < 0 || !out || (count && (!order || !names || !off || !tabs || !dmats)))
    return sats_fail(SATS_ERR_ARG, "sats_db_from_arrays: bad argument");
  std::unique_ptr<sats_db> db(new sats_db());
  db->tri_off.push_back(0);
  std::vector<uint8_t> tt;
  std::vector<float> td;
  for (int e = 0; e < count; e++) {
    int n = order[e];
    if (n < 1 || n > SATS_MAXDIM_EXT) return sats_fail(SATS_ERR_ARG, "entry %d has order %d outside 1..%d", e, n, SATS_MAXDIM_EXT);
    if (off[e] < 0) return sats_fail(SATS_ERR_ARG, "entry %d has a negative offset", e);
    tt.resize((size_t)n * (n + 1) / 2);
    td.resize(tt.size());
    const uint8_t *t = tabs + off[e];
    const float *d = dmats + off[e];
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        tt[(size_t)i * (i + 1) / 2 + j] = t[(size_t)i * n + j];
        td[(size_t)i * (i + 1) / 2 + j] = d[(size_t)i * n + j];
      }
    char nm[9] = {0};
    memcpy(nm, names + (size_t)e * 9, 8);
    if (int rc = check_codes("sats_db_from_arrays", nm, n, tt.data())) return rc;
    db->append(nm, n, tt.data(), td.data());
  }
  *out = db.release();
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" void sats_db_free(sats_db *db) { delete db; }
extern "C" int sats_db_count(const sats_db *db) { return db ? db->count() : 0; }
extern "C" int sats_db_order(const sats_db *db, int i) { return (db && i >= 0 && i < db->count()) ? db->order[i] : SATS_ERR_ARG; }
extern "C" const char *sats_db_name(const sats_db *db, int i) { return (db && i >= 0 && i < db->count()) ? db->name(i) : ""; }

extern "C" int sats_db_max_order(const sats_db *db)
{
  int m = 0;
  if (db) for (int n : db->order) m = std::max(m, n);
  return m;
}

extern "C" int sats_db_get(const sats_db *db, int e, uint8_t *tab, float *dmat)
{
  if (!db || e < 0 || e >= db->count()) return sats_fail(SATS_ERR_ARG, "sats_db_get: bad index %d", e);
  int n = db->order[e];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      if (tab) tab[(size_t)i * n + j] = db->code(e, i, j);
      if (dmat) dmat[(size_t)i * n + j] = db->dist(e, i, j);
    }
  return SATS_OK;
}

extern "C" int sats_db_find(const sats_db *db, const char *name)
{
  if (!db || !name) return sats_fail(SATS_ERR_ARG, "sats_db_find: null argument");
  for (int i = 0; i < db->count(); i++)
    if (!strcasecmp(name, db->name(i))) return i;
  return sats_fail(SATS_ERR_NOTFOUND, "ERROR: query %s not found", name);
}

extern "C" int sats_db_select(const sats_db *src, const int32_t *index, int count, sats_db **out)
try {
  if (!src || !out || count < 0 || (count && !index)) return sats_fail(SATS_ERR_ARG, "sats_db_select: bad argument");
  std::unique_ptr<sats_db> db(new sats_db());
  db->tri_off.push_back(0);
  for (int k = 0; k < count; k++) {
    int e = index[k];
    if (e < 0 || e >= src->count()) return sats_fail(SATS_ERR_ARG, "sats_db_select: bad index %d", e);
    db->append(src->name(e), src->order[e], src->tab.data() + src->tri_off[e], src->dmat.data() + src->tri_off[e]);
  }
  *out = db.release();
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" int sats_db_bootstrap(const sats_db *src, int count, uint64_t seed, int sort_by_order, sats_db **out)
try {
  if (!src || !out || count < 0 || src->count() == 0) return sats_fail(SATS_ERR_ARG, "sats_db_bootstrap: bad argument");
  uint64_t x = seed ? seed : 0x9E3779B97F4A7C15ull;
  std::vector<int32_t> pick((size_t)count);
  for (int k = 0; k < count; k++) {
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;              // xorshift64*
    uint64_t r = x * 0x2545F4914F6CDD1Dull;
    pick[k] = (int32_t)((r >> 33) % (uint64_t)src->count());
  }
  std::vector<int32_t> pos((size_t)count);
  std::iota(pos.begin(), pos.end(), 0);
  if (sort_by_order)
    std::stable_sort(pos.begin(), pos.end(), [&](int a, int b) { return src->order[pick[a]] < src->order[pick[b]]; });
  std::unique_ptr<sats_db> db(new sats_db());
  db->tri_off.push_back(0);
  char nm[16];
  for (int k = 0; k < count; k++) {
    int e = pick[pos[k]];
    snprintf(nm, sizeof nm, "s%06d", pos[k] % 1000000);
    db->append(nm, src->order[e], src->tab.data() + src->tri_off[e], src->dmat.data() + src->tri_off[e]);
  }
  *out = db.release();
  return SATS_OK;
}
SATS_CATCH_ALL

// ------------------------------------------------------------------------------------------ writers
// "%6.3f " of a distance without going through printf.  A float times 1000 is exact in double (24 + 10 significant bits),
// so nearbyint() of it is the round-half-even of the exact value -- what printf prints.  Anything that does not fit
// "dd.ddd" (negative, -0, >= 99.9995, non-finite) takes the printf route.
static inline void put_distance(std::string &out, float d)
{
  double v = std::isnan(d) ? 0.0 : (double)d;      // convdb2.py: NaN -> 0.000
  if (!std::signbit(v) && v < 99.9995) {
    const long k = std::lrint(v * 1000.0);         // default rounding mode: to nearest, ties to even
    if (k <= 99999) {
      char buf[7];
      const int ip = (int)(k / 1000), fp = (int)(k % 1000);
      buf[0] = ip >= 10 ? (char)('0' + ip / 10) : ' ';
      buf[1] = (char)('0' + ip % 10);
      buf[2] = '.';
      buf[3] = (char)('0' + fp / 100);
      buf[4] = (char)('0' + fp / 10 % 10);
      buf[5] = (char)('0' + fp % 10);
      buf[6] = ' ';
      out.append(buf, 7);
      return;
    }
  }
  char buf[64];
  int n = snprintf(buf, sizeof buf, "%6.3f ", v);
  out.append(buf, (size_t)n);
}

extern "C" int sats_db_write_ascii(const sats_db *db, const char *path)
try {
  if (!db || !path) return sats_fail(SATS_ERR_ARG, "sats_db_write_ascii: null argument");
  FILE *fp = fopen(path, "w");
  if (!fp) return sats_fail(SATS_ERR_IO, "cannot open %s for writing", path);
  static const char HI[] = "PROL?", LO[] = "EDST?";
  static const char *TY[] = {"e ", "xa", "xi", "xg"};
  std::string out;
  char head[64];
  for (int e = 0; e < db->count(); e++) {
    int n = db->order[e];
    out.clear();
    if (e) out.push_back('\n');
    out.append(head, (size_t)snprintf(head, sizeof head, "%6s %4d\n", db->name(e), n));
    for (int i = 0; i < n; i++) {
      for (int j = 0; j <= i; j++) {
        uint8_t v = db->code(e, i, j);
        if (i == j) { out.append(TY[v & 3], 2); out.push_back(' '); }
        else { out.push_back(HI[std::min(v >> 4, 4)]); out.push_back(LO[std::min(v & 15, 4)]); out.push_back(' '); }
      }
      out.push_back('\n');
    }
    for (int i = 0; i < n; i++) {
      for (int j = 0; j <= i; j++) put_distance(out, db->dist(e, i, j));
      out.push_back('\n');
    }
    if (fwrite(out.data(), 1, out.size(), fp) != out.size()) { fclose(fp); return sats_fail(SATS_ERR_IO, "write error on %s", path); }
  }
  if (fclose(fp)) return sats_fail(SATS_ERR_IO, "write error on %s", path);
  return SATS_OK;
}
SATS_CATCH_ALL

static const char PACKED_MAGIC[8] = {'S', 'A', 'T', 'S', 'D', 'B', '1', 0};

extern "C" int sats_db_write_packed(const sats_db *db, const char *path)
{
  if (!db || !path) return sats_fail(SATS_ERR_ARG, "sats_db_write_packed: null argument");
  FILE *fp = fopen(path, "wb");
  if (!fp) return sats_fail(SATS_ERR_IO, "cannot open %s for writing", path);
  uint32_t count = (uint32_t)db->count(), zero = 0;
  uint64_t cells = (uint64_t)db->tab.size();
  static const char pad[8] = {0};
  fwrite(PACKED_MAGIC, 1, 8, fp);
  fwrite(&count, 4, 1, fp); fwrite(&zero, 4, 1, fp); fwrite(&cells, 8, 1, fp);
  fwrite(db->order.data(), 4, count, fp);
  fwrite(db->names.data(), 1, (size_t)count * 9, fp);
  fwrite(pad, 1, (8 - (13 * (size_t)count) % 8) % 8, fp);
  fwrite(db->tab.data(), 1, cells, fp);
  fwrite(pad, 1, (8 - cells % 8) % 8, fp);
  fwrite(db->dmat.data(), 4, cells, fp);
  if (fclose(fp)) return sats_fail(SATS_ERR_IO, "write error on %s", path);
  return SATS_OK;
}

extern "C" int sats_db_read_packed(const char *path, sats_db **out)
try {
  if (!path || !out) return sats_fail(SATS_ERR_ARG, "sats_db_read_packed: null argument");
  std::string raw;
  int rc = slurp(path, &raw);
  if (rc) return rc;
  if (raw.size() < 24 || memcmp(raw.data(), PACKED_MAGIC, 8)) return sats_fail(SATS_ERR_PARSE, "%s is not a SATSDB1 file", path);
  uint32_t count; uint64_t cells;
  memcpy(&count, raw.data() + 8, 4);
  memcpy(&cells, raw.data() + 16, 8);
  // every field is bounded by the file size before any arithmetic on it (a corrupt header must not wrap `need`)
  if ((uint64_t)count > raw.size() / 13 || cells > raw.size() / 5) return sats_fail(SATS_ERR_PARSE, "%s is truncated", path);
  size_t pos = 24;
  size_t need = pos + 13 * (size_t)count + (8 - (13 * (size_t)count) % 8) % 8 + cells + (8 - cells % 8) % 8 + 4 * cells;
  if (raw.size() < need) return sats_fail(SATS_ERR_PARSE, "%s is truncated", path);
  std::unique_ptr<sats_db> db(new sats_db());
  db->order.resize(count);
  memcpy(db->order.data(), raw.data() + pos, 4 * (size_t)count); pos += 4 * (size_t)count;
  // the orders must account for exactly `cells` triangle cells before anything is copied by them
  db->tri_off.assign(1, 0);
  uint64_t sum = 0;
  for (uint32_t e = 0; e < count; e++) {
    int n = db->order[e];
    if (n < 1 || n > SATS_MAXDIM_EXT) return sats_fail(SATS_ERR_PARSE, "%s: entry %u has order %d", path, e, n);
    sum += (uint64_t)n * (n + 1) / 2;
    db->tri_off.push_back((int64_t)sum);
  }
  if (sum != cells) return sats_fail(SATS_ERR_PARSE, "%s: cell count mismatch", path);
  db->names.assign(raw.data() + pos, raw.data() + pos + 9 * (size_t)count); pos += 9 * (size_t)count;
  for (uint32_t e = 0; e < count; e++) db->names[(size_t)e * 9 + 8] = 0;
  pos += (8 - (13 * (size_t)count) % 8) % 8;
  db->tab.assign((const uint8_t *)raw.data() + pos, (const uint8_t *)raw.data() + pos + cells); pos += cells + (8 - cells % 8) % 8;
  db->dmat.resize(cells);
  memcpy(db->dmat.data(), raw.data() + pos, 4 * cells);
  for (uint32_t e = 0; e < count; e++)
    if (int crc = check_codes(path, db->name((int)e), db->order[e], db->tab.data() + db->tri_off[e])) return crc;
  *out = db.release();
  return SATS_OK;
}
SATS_CATCH_ALL

// ------------------------------------------------------------------------------------------ query construction
// scripts/pttableau.py:434-469 (angle_to_tabcode): the double quadrant encoding, every interval half-open on the left.
// Angles outside (-pi, pi] and NaN raise ValueError there; here they return SATS_ERR_ARG.
extern "C" int sats_tabcode_from_angle(double omega, char code[3])
{
  if (!code) return sats_fail(SATS_ERR_ARG, "sats_tabcode_from_angle: null argument");
  const double pi = M_PI;
  char a, b;
  if (omega > -pi / 4 && omega <= pi / 4) a = 'P';                                                       // parallel
  else if (omega > pi / 4 && omega <= 3 * pi / 4) a = 'R';                                               // crossing-right
  else if ((omega > 3 * pi / 4 && omega <= pi) || (omega > -pi && omega <= -3 * pi / 4)) a = 'O';        // antiparallel
  else if (omega > -3 * pi / 4 && omega <= -pi / 4) a = 'L';                                             // crossing-left
  else return sats_fail(SATS_ERR_ARG, "bad omega value %g", omega);
  if (omega > 0 && omega <= pi / 2) b = 'D';                                                             // dinner
  else if (omega > pi / 2 && omega <= pi) b = 'T';                                                       // tea
  else if (omega > -pi && omega <= -pi / 2) b = 'S';                                                     // supper
  else if (omega > -pi / 2 && omega <= 0) b = 'E';                                                       // elevenses
  else return sats_fail(SATS_ERR_ARG, "bad omega value %g", omega);
  code[0] = a; code[1] = b; code[2] = 0;
  return SATS_OK;
}

namespace {
struct V3 { double x, y, z; };
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 axpy(V3 p, double s, V3 d) { return {p.x + s * d.x, p.y + s * d.y, p.z + s * d.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 unit(V3 a) { const double n = std::sqrt(dot(a, a)); return {a.x / n, a.y / n, a.z / n}; }
inline bool tiny(V3 a, double eps) { return std::fabs(a.x) < eps && std::fabs(a.y) < eps && std::fabs(a.z) < eps; }
}  // namespace

// scripts/ptnode.py:752-880 (PTNode.relative_angle) with scripts/geometry.py:18-79 (LineLineIntersect, Bourke's shortest
// line between two lines): the interaxial angle omega in (-pi, pi] of the axis of `self` (centroid c_self, direction cosines
// d_self) and the axis of the other SSE, looking along their common perpendicular.  Returns 1 (and leaves *omega alone) where
// the reference returns None: degenerate axis or no unique common perpendicular (parallel axes).
extern "C" int sats_relative_angle(const double c_self[3], const double d_self[3], const double c_other[3],
                                   const double d_other[3], double *omega)
{
  if (!c_self || !d_self || !c_other || !d_other || !omega) return sats_fail(SATS_ERR_ARG, "sats_relative_angle: null argument");
  const double ALPHA = 100.0, EPS = 1.0e-08;            // ptnode.py:43, geometry.py:47
  const V3 cs{c_self[0], c_self[1], c_self[2]}, ds{d_self[0], d_self[1], d_self[2]};
  const V3 co{c_other[0], c_other[1], c_other[2]}, dother{d_other[0], d_other[1], d_other[2]};
  const V3 p1 = co, p2 = axpy(co, ALPHA, dother);          // pa: a second point on the other SSE's axis
  const V3 p3 = cs, p4 = axpy(cs, ALPHA, ds);          // pd: a second point on this SSE's axis
  const V3 p13 = p1 - p3, p43 = p4 - p3;
  if (tiny(p43, EPS)) return 1;
  const V3 p21 = p2 - p1;
  if (tiny(p21, EPS)) return 1;
  const double d1343 = dot(p13, p43), d4321 = dot(p43, p21), d1321 = dot(p13, p21), d4343 = dot(p43, p43), d2121 = dot(p21, p21);
  const double denom = d2121 * d4343 - d4321 * d4321;
  if (std::fabs(denom) < EPS) return 1;
  const double numer = d1343 * d4321 - d1321 * d4343;
  const double mua = numer / denom, mub = (d1343 + d4321 * mua) / d4343;
  const V3 pb = axpy(p1, mua, p21), pc = axpy(p3, mub, p43);
  const V3 v1 = pb - p2, v2 = pc - pb, v3 = p4 - pc;
  const V3 n1 = unit(cross(v1, v2)), n2 = unit(cross(v2, v3));
  double c = dot(n1, n2);
  if (1.0 < c) c = 1.0;             // python's min(c, 1) / max(c, -1): a NaN passes through
  if (-1.0 > c) c = -1.0;
  double om = std::acos(c);
  if (dot(v2, cross(n1, n2)) < 0) om = -om;
  *omega = om;
  return SATS_OK;
}

// A structure from SSE axes: compute_tableau (scripts/pttableau.py:470-520, use_hk = False as in the search databases) and
// compute_sse_midpoint_dist_matrix (scripts/ptdistmatrix.py:1014-1066), then the writer's conventions -- "%6.3f" text that the
// search program reads back with strtof (scripts/convdb2.py:213-231), NaN -> 0.000, distances above 99.9 A clamped to 99.9
// (scripts/pytableaucreate.py:114-116; wider values break the 7-column format).
static int build_structure_impl(const char *name, int n, const uint8_t *sse_type, const double *centroid, const double *dircos,
                                const uint8_t *has_axis, sats_db **out)
{
  if (!name || !sse_type || !centroid || !dircos || !out) return sats_fail(SATS_ERR_ARG, "sats_build_structure: null argument");
  if (n < 1 || n > SATS_MAXDIM) return sats_fail(SATS_ERR_ARG, "sats_build_structure: order %d outside 1..%d", n, SATS_MAXDIM);
  std::vector<uint8_t> tab((size_t)n * (n + 1) / 2);
  std::vector<float> dm(tab.size());
  auto as_written = [](double d) {
    if (std::isnan(d)) d = 0.0;
    if (d > 99.9) d = 99.9;
    char buf[64];
    snprintf(buf, sizeof buf, "%6.3f", d);
    return strtof(buf, nullptr);
  };
  for (int i = 0; i < n; i++) {
    if (sse_type[i] > 3) return sats_fail(SATS_ERR_ARG, "sats_build_structure: SSE %d has type %d (0 strand, 1 alpha, 2 pi, 3 3-10)", i, sse_type[i]);
    tab[(size_t)i * (i + 1) / 2 + i] = sse_type[i];
    dm[(size_t)i * (i + 1) / 2 + i] = (float)sse_type[i];
  }
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      double omega = 0.0;
      int code = 0x44;                                         // "??": no angle (the reference leaves the entry unset)
      const bool axes = !has_axis || (has_axis[i] && has_axis[j]);        // an SSE without an axis: fit_axis() returned None
      const int rc = axes ? sats_relative_angle(centroid + 3 * i, dircos + 3 * i, centroid + 3 * j, dircos + 3 * j, &omega) : 1;
      if (rc < 0) return rc;
      if (rc == 0) {
        char c2[3];
        if (sats_tabcode_from_angle(omega, c2) != SATS_OK) {
          fprintf(stderr, "WARNING: catch bad tableau angle, seting Parallel (%d,%d)\n", i, j);      // pttableau.py:497-499
          c2[0] = 'P'; c2[1] = 'E';
        }
        if (tableau_code(c2[0], c2[1], &code)) return SATS_ERR_PARSE;
      }
      const double dx = centroid[3 * i] - centroid[3 * j], dy = centroid[3 * i + 1] - centroid[3 * j + 1],
                   dz = centroid[3 * i + 2] - centroid[3 * j + 2];
      tab[(size_t)j * (j + 1) / 2 + i] = (uint8_t)code;
      // calc_sse_sse_midpoint_dist returns None without both axes; the matrix then holds NaN, written as 0.000
      dm[(size_t)j * (j + 1) / 2 + i] = as_written(axes ? std::sqrt(dx * dx + dy * dy + dz * dz) : NAN);
    }
  std::unique_ptr<sats_db> db(new sats_db());
  db->append(name, n, tab.data(), dm.data());
  *out = db.release();
  return SATS_OK;
}

extern "C" int sats_build_structure(const char *name, int n, const uint8_t *sse_type, const double *centroid,
                                    const double *dircos, sats_db **out)
try {
  return build_structure_impl(name, n, sse_type, centroid, dircos, nullptr, out);
}
SATS_CATCH_ALL

// scripts/ptnode.py:1113-1292 (PTNodeHelix.fit_axis) and :1846-1990 (PTNodeStrand.fit_axis): the axis of an SSE as a total
// least squares line through points derived from its C-alpha trace -- helices: the midpoints of the planes of consecutive
// C-alpha triples; strands: the midpoints of consecutive C-alpha pairs (to take out the pleat), centred on the C-alpha
// centroid.  The direction is the first right singular vector of the centred points (here: the dominant eigenvector of their
// 3 x 3 scatter matrix, by Jacobi rotations), oriented from the N- to the C-terminus.  Short SSEs fall back to the line through
// two midpoints / two atoms; returns 1 where the reference returns None (helix of < 3 residues, strand of 1).
namespace {
V3 dominant_direction(const std::vector<V3> &pts)
{
  double a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (const V3 &p : pts) {
    const double c[3] = {p.x, p.y, p.z};
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) a[r][q] += c[r] * c[q];
  }
  for (int sweep = 0; sweep < 64; sweep++) {                    // cyclic Jacobi on the symmetric 3 x 3 scatter matrix
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; k++) {                            // A <- A J
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - sn * akq;
          a[k][q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {                            // A <- J^T A
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - sn * aqk;
          a[q][k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {                            // V <- V J
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - sn * vkq;
          v[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int k = 1; k < 3; k++) if (a[k][k] > a[best][best]) best = k;
  return unit(V3{v[0][best], v[1][best], v[2][best]});
}
}  // namespace

extern "C" int sats_fit_axis(int sse_type, int n_res, const double *ca_xyz, double dircos[3], double centroid[3])
try {
  if (!ca_xyz || !dircos || !centroid || n_res < 0 || sse_type < 0 || sse_type > 3) return sats_fail(SATS_ERR_ARG, "sats_fit_axis: bad argument");
  auto ca = [&](int i) { return V3{ca_xyz[3 * i], ca_xyz[3 * i + 1], ca_xyz[3 * i + 2]}; };
  auto mid = [&](V3 a, V3 b) { return V3{(b.x - a.x) / 2 + a.x, (b.y - a.y) / 2 + a.y, (b.z - a.z) / 2 + a.z}; };
  auto mean = [&](const std::vector<V3> &p) {
    V3 c{0, 0, 0};
    for (const V3 &q : p) { c.x += q.x; c.y += q.y; c.z += q.z; }
    return V3{c.x / (double)p.size(), c.y / (double)p.size(), c.z / (double)p.size()};
  };
  std::vector<V3> pts;
  V3 cen{0, 0, 0}, dir{0, 0, 0}, nterm{0, 0, 0}, cterm{0, 0, 0};
  bool fitted = false;
  if (sse_type != 0) {                                            // helix (alpha, pi, 3-10)
    if (n_res < 3) return 1;
    if (n_res == 3) {
      const V3 mp1 = mid(ca(0), ca(1)), mp2 = mid(ca(1), ca(2));
      cen = V3{(mp1.x + mp2.x) / 2, (mp1.y + mp2.y) / 2, (mp1.z + mp2.z) / 2};
      dir = unit(mp2 - mp1);
    } else {
      for (int i = 1; i + 1 < n_res; i++) {                       // midpoint of the plane of C-alphas i-1, i, i+1
        const V3 v1 = ca(i - 1) - ca(i), v2 = ca(i + 1) - ca(i);
        pts.push_back(V3{ca(i).x + (v1.x + v2.x) / 2, ca(i).y + (v1.y + v2.y) / 2, ca(i).z + (v1.z + v2.z) / 2});
      }
      cen = mean(pts);
      nterm = pts.front(); cterm = pts.back();
      for (V3 &q : pts) q = q - cen;
      fitted = true;
    }
  } else {                                                        // strand
    if (n_res < 2) return 1;
    std::vector<V3> atoms;
    for (int i = 0; i < n_res; i++) atoms.push_back(ca(i));
    cen = mean(atoms);
    if (n_res == 2) dir = unit(ca(1) - ca(0));
    else if (n_res == 3) dir = unit(mid(ca(1), ca(2)) - mid(ca(0), ca(1)));
    else {
      for (int i = 0; i + 1 < n_res; i++) pts.push_back(mid(ca(i), ca(i + 1)) - cen);
      nterm = ca(0); cterm = ca(n_res - 1);
      fitted = true;
    }
  }
  if (fitted) {
    dir = dominant_direction(pts);
    // orientation: N- to C-terminus (the reference projects both ends onto the line and tests the angle for pi)
    if (dot(cterm - nterm, dir) < 0) dir = V3{-dir.x, -dir.y, -dir.z};
  }
  dircos[0] = dir.x; dircos[1] = dir.y; dircos[2] = dir.z;
  centroid[0] = cen.x; centroid[1] = cen.y; centroid[2] = cen.z;
  return SATS_OK;
}
SATS_CATCH_ALL

// C-alpha traces of the SSEs -> searchable structure: sats_fit_axis per SSE, then sats_build_structure; an SSE whose axis cannot
// be fitted keeps "??" codes and 0.000 distances against everything, as the reference's None propagates.
extern "C" int sats_build_structure_from_ca(const char *name, int n, const uint8_t *sse_type, const int32_t *n_res,
                                            const double *ca_xyz, sats_db **out)
try {
  if (!name || !sse_type || !n_res || !ca_xyz || !out) return sats_fail(SATS_ERR_ARG, "sats_build_structure_from_ca: null argument");
  if (n < 1 || n > SATS_MAXDIM) return sats_fail(SATS_ERR_ARG, "sats_build_structure_from_ca: order %d outside 1..%d", n, SATS_MAXDIM);
  std::vector<double> cen(3 * (size_t)n, 0.0), dir(3 * (size_t)n, 0.0);
  std::vector<uint8_t> has((size_t)n, 0);
  size_t at = 0;
  for (int i = 0; i < n; i++) {
    if (n_res[i] < 0) return sats_fail(SATS_ERR_ARG, "sats_build_structure_from_ca: SSE %d has %d residues", i, n_res[i]);
    if (sse_type[i] > 3) return sats_fail(SATS_ERR_ARG, "sats_build_structure_from_ca: SSE %d has type %d", i, sse_type[i]);
    const int rc = sats_fit_axis(sse_type[i], n_res[i], ca_xyz + 3 * at, &dir[3 * (size_t)i], &cen[3 * (size_t)i]);
    if (rc < 0) return rc;
    has[(size_t)i] = rc == 0;
    if (rc == 1) fprintf(stderr, "WARNING: SSE %d has only %d residues, cannot fit axis\n", i, n_res[i]);
    at += (size_t)n_res[i];
  }
  return build_structure_impl(name, n, sse_type, cen.data(), dir.data(), has.data(), out);
}
SATS_CATCH_ALL

// ------------------------------------------------------------------------------------------ statistics
extern "C" {
const double sats_gumbel_a = 0.3780327676087335;
const double sats_gumbel_b = 0.3582596175507505;
}
static const double kEulerGamma = 0.5772156649015328606;

extern "C" double sats_norm2(int score, int size1, int size2) { return 2.0 * score / ((double)(size1 + size2)); }

extern "C" double sats_z_gumbel(int x, double a, double b)
{
  double mu = a + b * kEulerGamma;
  double sigma = (M_PI / sqrt(6.0)) * b;
  return (x - mu) / sigma;
}

extern "C" double sats_pv_gumbel(double z) { return 1 - exp(-exp(-((M_PI / sqrt(6.0)) * z + kEulerGamma))); }

// z is a non-decreasing step function of the raw score for fixed sizes (norm2 = 2 score / (n1 + n2), truncated to int at the
// z_gumbel call like the reference does): bisect for the first score that passes.  |score| <= 2 C(111, 2) = 12210.
extern "C" int32_t sats_score_threshold(double z_min, int size1, int size2)
{
  auto pass = [&](int sc) { return sats_z_gumbel((int)sats_norm2(sc, size1, size2), sats_gumbel_a, sats_gumbel_b) >= z_min; };
  int lo = -16384, hi = 16384;
  if (!(z_min == z_min) || size1 + size2 < 1 || !pass(hi)) return INT32_MAX;
  while (lo < hi) {
    int mid = lo + (hi - lo) / 2;
    if (pass(mid)) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

extern "C" size_t sats_format_block(char *buf, size_t cap, const char *query_id, int query_order, const char *dbfile,
                                    int lorder, int lsoln, const sats_db *db, const int32_t *index, int count,
                                    const int32_t *scores, const int32_t *maps)
{
  size_t used = 0;
  auto emit = [&](const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(used < cap ? buf + used : nullptr, used < cap ? cap - used : 0, fmt, ap);
    va_end(ap);
    if (n > 0) used += (size_t)n;
  };
  auto put = [&](const char *src, size_t n) {          // appends like emit() does: counts everything, writes what fits
    if (used < cap) memcpy(buf + used, src, std::min(n, cap - used - 1)), buf[std::min(used + n, cap - 1)] = 0;
    used += n;
  };
  auto put_name = [&](const char *name) {               // "%-8s": left-justified, padded to 8, never truncated
    char nm[16] = "        ";
    size_t n = strlen(name);
    if (n > 8) { put(name, n); return; }
    memcpy(nm, name, n);
    put(nm, 8);
  };
  emit("# cudaSaTabsearch LTYPE = %c LORDER = %c LSOLN = %c\n", 'T', lorder ? 'T' : 'F', lsoln ? 'T' : 'F');
  emit("# QUERY ID = %-8s\n", query_id);
  emit("# DBFILE = %-80s\n", dbfile);
  // The four numeric columns depend on (score, structure order) only -- a few thousand distinct pairs in a block of any
  // size -- so each distinct tail " %d %g %g %g\n" goes through printf once and is copied afterwards (a 200-query run
  // prints 2.9 M rows).  Key: order in the low 8 bits, score above.
  struct Tail { char text[80]; uint8_t len; };
  std::vector<std::pair<int64_t, Tail>> cache;          // open addressing, power-of-two size
  cache.assign(1u << 14, std::make_pair((int64_t)-1, Tail()));
  size_t filled = 0;
  for (int k = 0; k < count; k++) {
    int e = index ? index[k] : k;
    const int64_t key = ((int64_t)(uint32_t)scores[e] << 8) | (int64_t)(db->order[e] & 255);
    size_t h = (size_t)((uint64_t)key * 0x9E3779B97F4A7C15ull >> 40) & (cache.size() - 1);
    while (cache[h].first != key && cache[h].first != -1) h = (h + 1) & (cache.size() - 1);
    if (cache[h].first != key) {
      double n2s = sats_norm2(scores[e], query_order, db->order[e]);
      double z = sats_z_gumbel((int)n2s, sats_gumbel_a, sats_gumbel_b);   // implicit double -> int at the call, as in the reference
      double pv = sats_pv_gumbel(z);
      Tail t;
      t.len = (uint8_t)snprintf(t.text, sizeof t.text, " %d %g %g %g\n", scores[e], n2s, z, pv);
      if (filled * 2 < cache.size()) { cache[h] = std::make_pair(key, t); filled++; }
      put_name(db->name(e));
      put(t.text, t.len);
    } else {
      put_name(db->name(e));
      put(cache[h].second.text, cache[h].second.len);
    }
    if (lsoln && maps)
      for (int i = 0; i < query_order; i++) {
        int j = maps[(size_t)e * SATS_MAP_STRIDE + i];
        if (j >= 0) emit("%3d %3d\n", i + 1, j + 1);
      }
  }
  return used;
}

// ------------------------------------------------------------------------------------------ sharding
extern "C" int sats_partition(const sats_db *db, int shard_count, int32_t *owner)
try {
  if (!db || !owner || shard_count < 1) return sats_fail(SATS_ERR_ARG, "sats_partition: bad argument");
  int n = db->count();
  std::vector<int32_t> idx((size_t)n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return db->order[a] > db->order[b]; });
  std::vector<double> load((size_t)shard_count, 0.0);
  for (int k = 0; k < n; k++) {     // longest-processing-time-first over the size-sorted list
    int best = 0;
    for (int s = 1; s < shard_count; s++)
      if (load[s] < load[best]) best = s;
    owner[idx[k]] = best;
    load[best] += sats_entry_cost(db->order[idx[k]]);
  }
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" void sats_params_default(sats_params *p)
{
  if (!p) return;
  memset(p, 0, sizeof *p);
  p->lorder = 1;
  p->lsoln = 0;
  p->restarts = SATS_DEFAULT_MAXSTART;
  p->rng_mode = SATS_RNG_PHILOX;
  p->accept_mode = SATS_ACCEPT_HOST_TABLE;
  p->pool = SATS_POOL_ALL;
  p->pool_threshold = SATS_MAXDIM_GPU;
  p->seed = SATS_REF_SEED;
}

// ------------------------------------------------------------------------------------------ result reader
struct sats_results {
  struct Row { char name[9]; int32_t score; double norm2, z, p; std::vector<int32_t> pairs; };
  struct Block { std::string query, dbfile; int flags[3] = {1, 1, 0}; std::vector<Row> rows; };
  std::vector<Block> blocks;
};

extern "C" int sats_results_parse(const char *text, size_t len, sats_results **out)
try {
  if (!text || !out) return sats_fail(SATS_ERR_ARG, "sats_results_parse: null argument");
  std::unique_ptr<sats_results> r(new sats_results());
  Cursor c{text, text + len};
  const char *b, *e;
  int lineno = 0;
  while (c.line(b, e)) {
    lineno++;
    std::string ln(b, e);
    while (!ln.empty() && (ln.back() == '\r' || ln.back() == ' ')) ln.pop_back();
    if (ln.empty()) continue;
    if (ln[0] == '#') {
      char f[3];
      if (sscanf(ln.c_str(), "# cudaSaTabsearch LTYPE = %c LORDER = %c LSOLN = %c", &f[0], &f[1], &f[2]) == 3) {
        r->blocks.emplace_back();
        for (int i = 0; i < 3; i++) r->blocks.back().flags[i] = f[i] == 'T';
      } else if (!r->blocks.empty() && ln.compare(0, 13, "# QUERY ID = ") == 0) {
        r->blocks.back().query = ln.substr(13);
      } else if (!r->blocks.empty() && ln.compare(0, 11, "# DBFILE = ") == 0) {
        r->blocks.back().dbfile = ln.substr(11);
      }
      continue;
    }
    if (r->blocks.empty()) r->blocks.emplace_back();       // headerless output (e.g. filtered through grep -v '#')
    sats_results::Row row;
    char nm[64];
    int a, bb;
    char extra;
    if (sscanf(ln.c_str(), "%63s %d %lf %lf %lf", nm, &row.score, &row.norm2, &row.z, &row.p) == 5) {
      memset(row.name, 0, 9);
      strncpy(row.name, nm, 8);
      r->blocks.back().rows.push_back(row);
    } else if (sscanf(ln.c_str(), "%d %d %c", &a, &bb, &extra) == 2 && !r->blocks.back().rows.empty()) {
      r->blocks.back().rows.back().pairs.push_back(a);
      r->blocks.back().rows.back().pairs.push_back(bb);
    } else {
      return sats_fail(SATS_ERR_PARSE, "result line %d not understood: %.60s", lineno, ln.c_str());
    }
  }
  *out = r.release();
  return SATS_OK;
}
SATS_CATCH_ALL

extern "C" void sats_results_free(sats_results *r) { delete r; }
extern "C" int sats_results_blocks(const sats_results *r) { return r ? (int)r->blocks.size() : 0; }
static const sats_results::Block *blk(const sats_results *r, int b) { return (r && b >= 0 && b < (int)r->blocks.size()) ? &r->blocks[b] : nullptr; }
extern "C" const char *sats_results_query(const sats_results *r, int b) { return blk(r, b) ? blk(r, b)->query.c_str() : ""; }
extern "C" const char *sats_results_dbfile(const sats_results *r, int b) { return blk(r, b) ? blk(r, b)->dbfile.c_str() : ""; }
extern "C" int sats_results_flags(const sats_results *r, int b, int flags_tf[3])
{
  if (!blk(r, b) || !flags_tf) return sats_fail(SATS_ERR_ARG, "sats_results_flags: bad block %d", b);
  for (int i = 0; i < 3; i++) flags_tf[i] = blk(r, b)->flags[i];
  return SATS_OK;
}
extern "C" int sats_results_rows(const sats_results *r, int b) { return blk(r, b) ? (int)blk(r, b)->rows.size() : SATS_ERR_ARG; }
extern "C" int sats_results_row(const sats_results *r, int b, int row, char name[9], int32_t *score, double *norm2, double *z,
                                double *pvalue)
{
  const sats_results::Block *k = blk(r, b);
  if (!k || row < 0 || row >= (int)k->rows.size()) return sats_fail(SATS_ERR_ARG, "sats_results_row: bad block/row %d/%d", b, row);
  const sats_results::Row &x = k->rows[row];
  if (name) memcpy(name, x.name, 9);
  if (score) *score = x.score;
  if (norm2) *norm2 = x.norm2;
  if (z) *z = x.z;
  if (pvalue) *pvalue = x.p;
  return SATS_OK;
}
extern "C" int sats_results_map(const sats_results *r, int b, int row, int32_t *pairs, int cap)
{
  const sats_results::Block *k = blk(r, b);
  if (!k || row < 0 || row >= (int)k->rows.size()) return sats_fail(SATS_ERR_ARG, "sats_results_map: bad block/row %d/%d", b, row);
  const std::vector<int32_t> &p = k->rows[row].pairs;
  int n = (int)p.size() / 2;
  for (int i = 0; i < n && i < cap && pairs; i++) { pairs[2 * i] = p[2 * i]; pairs[2 * i + 1] = p[2 * i + 1]; }
  return n;
}

extern "C" double sats_roc_auc(const double *score, const uint8_t *positive, int n)
{
  if (!score || !positive || n < 2) return NAN;
  std::vector<int> idx((size_t)n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return score[a] < score[b]; });
  // rank-sum with mid-ranks for ties
  double ranksum = 0.0;
  long long npos = 0;
  for (int i = 0; i < n;) {
    int j = i;
    while (j < n && score[idx[j]] == score[idx[i]]) j++;
    double midrank = 0.5 * ((i + 1) + j);
    for (int t = i; t < j; t++) if (positive[idx[t]]) { ranksum += midrank; npos++; }
    i = j;
  }
  long long nneg = n - npos;
  if (npos == 0 || nneg == 0) return NAN;
  return (ranksum - 0.5 * (double)npos * (double)(npos + 1)) / ((double)npos * (double)nneg);
}
