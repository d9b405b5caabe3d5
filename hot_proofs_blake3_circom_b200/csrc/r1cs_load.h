// r1cs_load.h -- host side of b3w_r1cs_load: parse an iden3 `.r1cs` file (binary format v1: magic "r1cs", sections
// 1 header / 2 constraints / 3 wire->label map; what circom writes next to the wasm and what snarkjs, circom_tester and
// circom-scotia read -- rust_fold/src/blake3_circuit.rs:71-81) and group its rows into the shape classes the device
// evaluator works on (r1cs.cuh): rows with the same (nA, nB, nC) and the same small coefficients share a coefficient
// vector; small groups are merged into per-row-coefficient classes; rows whose coefficients do not fit 56 bits (e.g.
// 2^-31 mod p after circom's O2 substitution) keep full field-element coefficients (BIGCOEF) and are evaluated in Fr.
// Also here: fp_compile(), which turns a row list (from a file or from the built-in tables) into the compiled program of
// kernels_r1cs_fast.cuh; only the rows it does not take go through the class grouping.
// Included by blake3wit.cu only.
#pragma once
#include <map>
#include <vector>
#include <stdint.h>

struct r1cs_host_set {
  std::vector<r1cs_class_dev> cls;
  std::vector<int64_t> lo, hi;          // small coefficients (two's complement 128-bit)
  std::vector<fr_t> coef_fr;            // BIGCOEF coefficients
  std::vector<uint32_t> terms, row_ids;
  uint32_t rows = 0, n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0;
};

namespace r1cs_load_detail {
struct term { uint32_t wire; bool small; __int128 c; fr_t f; };      // small: the coefficient is the signed integer c (|c| < 2^120)
struct row { std::vector<term> part[3]; uint32_t id; bool big; };

static bool rd32(const uint8_t *p, size_t len, size_t &pos, uint32_t &v) {
  if (pos + 4 > len) return false;
  memcpy(&v, p + pos, 4);
  pos += 4;
  return true;
}
static bool rd64(const uint8_t *p, size_t len, size_t &pos, uint64_t &v) {
  if (pos + 8 > len) return false;
  memcpy(&v, p + pos, 8);
  pos += 8;
  return true;
}
}  // namespace r1cs_load_detail

// returns 0 or a B3W_ERR_* code with the text in `err`.  Rows come back in file order with R.id = constraint index.
static int r1cs_parse_rows(const uint8_t *data, size_t len, const uint8_t prime[32], uint32_t ws, std::vector<r1cs_load_detail::row> &rows,
                           r1cs_host_set &out, std::string &err) {
  using namespace r1cs_load_detail;
  size_t pos = 0;
  uint32_t version, nsec;
  if (len < 12 || memcmp(data, "r1cs", 4) != 0) { err = "not an r1cs file (magic)"; return B3W_ERR_INVALID; }
  pos = 4;
  if (!rd32(data, len, pos, version) || !rd32(data, len, pos, nsec) || version != 1) { err = "unsupported r1cs version"; return B3W_ERR_UNSUPPORTED; }
  size_t sec_off[4] = {0, 0, 0, 0}, sec_len[4] = {0, 0, 0, 0};
  for (uint32_t i = 0; i < nsec; i++) {
    uint32_t type;
    uint64_t sz;
    if (!rd32(data, len, pos, type) || !rd64(data, len, pos, sz) || sz > len - pos) { err = "truncated section table"; return B3W_ERR_INVALID; }
    if (type >= 1 && type <= 3) { sec_off[type] = pos; sec_len[type] = (size_t)sz; }
    pos += (size_t)sz;
  }
  if (!sec_off[1] || !sec_off[2]) { err = "header or constraint section missing"; return B3W_ERR_INVALID; }
  // header
  pos = sec_off[1];
  const size_t hend = sec_off[1] + sec_len[1];
  uint32_t fs, n_wires, n_pub_out, n_pub_in, n_prv_in, m;
  uint64_t n_labels;
  if (!rd32(data, hend, pos, fs) || fs != 32 || pos + 32 > hend) { err = "field size is not 32 bytes"; return B3W_ERR_UNSUPPORTED; }
  if (memcmp(data + pos, prime, 32) != 0) { err = "the file's prime is not this circuit's prime"; return B3W_ERR_INVALID; }
  pos += 32;
  if (!rd32(data, hend, pos, n_wires) || !rd32(data, hend, pos, n_pub_out) || !rd32(data, hend, pos, n_pub_in) ||
      !rd32(data, hend, pos, n_prv_in) || !rd64(data, hend, pos, n_labels) || !rd32(data, hend, pos, m)) { err = "truncated header"; return B3W_ERR_INVALID; }
  if (n_wires != ws) { err = "the file has " + std::to_string(n_wires) + " wires, this circuit's witness has " + std::to_string(ws); return B3W_ERR_INVALID; }
  out.n_wires = n_wires; out.n_pub_out = n_pub_out; out.n_pub_in = n_pub_in; out.n_prv_in = n_prv_in;
  fr_t P;
  memcpy(P.l, prime, 32);
  // constraints
  pos = sec_off[2];
  const size_t cend = sec_off[2] + sec_len[2];
  if ((uint64_t)m * 12 > (uint64_t)(cend - pos)) { err = "the header announces more constraints than the file holds"; return B3W_ERR_INVALID; }
  rows.assign(m, row());
  for (uint32_t i = 0; i < m; i++) {
    row &R = rows[i];
    R.id = i;
    R.big = false;
    for (int part = 0; part < 3; part++) {
      uint32_t k;
      if (!rd32(data, cend, pos, k) || (size_t)k * 36 > cend - pos) { err = "truncated constraint " + std::to_string(i); return B3W_ERR_INVALID; }
      if (k > 65535) { err = "constraint " + std::to_string(i) + " has " + std::to_string(k) + " terms in one linear combination"; return B3W_ERR_UNSUPPORTED; }
      R.part[part].resize(k);
      for (uint32_t j = 0; j < k; j++) {
        term &T = R.part[part][j];
        memcpy(&T.wire, data + pos, 4);
        memcpy(T.f.l, data + pos + 4, 32);
        pos += 36;
        if (T.wire >= n_wires) { err = "constraint " + std::to_string(i) + " refers to wire " + std::to_string(T.wire); return B3W_ERR_INVALID; }
        if (fr_gte(T.f, P)) { err = "constraint " + std::to_string(i) + " has a coefficient >= p"; return B3W_ERR_INVALID; }
        // small signed form: c < 2^120 or p - c < 2^120 (the class grouping keeps only < 2^56 as "small")
        fr_t neg;
        fr_raw_sub(neg, P, T.f);
        auto fits = [](const fr_t &x) { return (x.l[4] | x.l[5] | x.l[6] | x.l[7]) == 0 && (x.l[3] >> 24) == 0; };
        auto val = [](const fr_t &x) {
          return (__int128)(((unsigned __int128)(((uint64_t)x.l[3] << 32) | x.l[2]) << 64) | (unsigned __int128)(((uint64_t)x.l[1] << 32) | x.l[0]));
        };
        if (fits(T.f)) { T.small = true; T.c = val(T.f); }
        else if (fits(neg)) { T.small = true; T.c = -val(neg); }
        else { T.small = false; T.c = 0; }
      }
    }
    if (R.part[0].empty() || R.part[1].empty()) { R.part[0].clear(); R.part[1].clear(); }      // 0 * B = C  <=>  C = 0
  }
  return B3W_OK;
}

// Group rows into the shape classes of the general evaluator (r1cs_rows.cuh).  `rows` is consumed (terms get sorted).
static void r1cs_group(std::vector<r1cs_load_detail::row> &rows, r1cs_host_set &out) {
  using namespace r1cs_load_detail;
  const __int128 lim56 = (__int128)1 << 56;
  for (row &R : rows) {
    R.big = false;
    for (int part = 0; part < 3; part++) {
      for (const term &T : R.part[part]) R.big = R.big || !T.small || T.c >= lim56 || T.c <= -lim56;
      if (R.part[part].size() > 64) R.big = true;          // keeps the 128-bit accumulators of the integer path exact
    }
    for (int part = 0; part < 3; part++)
      std::sort(R.part[part].begin(), R.part[part].end(), [&](const term &a, const term &b) {
        if (!R.big && a.c != b.c) return a.c < b.c;
        return a.wire < b.wire;
      });
  }
  const uint32_t m = (uint32_t)rows.size();
  typedef std::vector<long long> key_t;                    // nA, nB, nC, big?, then (hi, lo) of every coefficient
  std::map<key_t, std::vector<uint32_t>> groups;
  for (uint32_t i = 0; i < m; i++) {
    const row &R = rows[i];
    key_t k = {(long long)R.part[0].size(), (long long)R.part[1].size(), (long long)R.part[2].size(), R.big ? 1 : 0};
    if (!R.big)
      for (int part = 0; part < 3; part++)
        for (const term &T : R.part[part]) { k.push_back((long long)(T.c >> 64)); k.push_back((long long)(uint64_t)T.c); }
    groups[k].push_back(i);
  }
  std::map<key_t, std::vector<uint32_t>> merged;           // small groups -> per-row-coefficient classes by (nA, nB, nC, big)
  auto emit = [&](const std::vector<uint32_t> &members, bool rowcoef, bool big) {
    const row &R0 = rows[members[0]];
    r1cs_class_dev c;
    c.nA = (uint16_t)R0.part[0].size(); c.nB = (uint16_t)R0.part[1].size(); c.nC = (uint16_t)R0.part[2].size();
    c.flags = (uint16_t)(R1CS_FLAG_WIDE | (rowcoef ? R1CS_FLAG_ROWCOEF : 0) | (big ? R1CS_FLAG_BIGCOEF : 0));
    c.count = (uint32_t)members.size();
    c.coef_off = (uint32_t)(big ? out.coef_fr.size() : out.lo.size());
    c.term_off = (uint32_t)out.terms.size();
    c.row_off = out.rows;
    const uint32_t nt = c.nA + c.nB + c.nC;
    for (uint32_t t = 0; t < nt; t++) {
      const int part = t < c.nA ? 0 : t < (uint32_t)c.nA + c.nB ? 1 : 2;
      const uint32_t j = part == 0 ? t : part == 1 ? t - c.nA : t - c.nA - c.nB;
      for (uint32_t r : members) out.terms.push_back(rows[r].part[part][j].wire);
      if (rowcoef) {
        for (uint32_t r : members) {
          const term &T = rows[r].part[part][j];
          if (big) out.coef_fr.push_back(T.f);
          else { out.lo.push_back((int64_t)(uint64_t)T.c); out.hi.push_back((int64_t)(T.c >> 64)); }
        }
      } else {
        const term &T = R0.part[part][j];
        out.lo.push_back((int64_t)(uint64_t)T.c);
        out.hi.push_back((int64_t)(T.c >> 64));
      }
    }
    for (uint32_t r : members) out.row_ids.push_back(rows[r].id);
    out.rows += c.count;
    out.cls.push_back(c);
  };
  for (auto &g : groups) {
    const bool big = g.first[3] != 0;
    if (big || g.second.size() < 16) {
      key_t k(g.first.begin(), g.first.begin() + 4);
      auto &v = merged[k];
      v.insert(v.end(), g.second.begin(), g.second.end());
    } else {
      emit(g.second, false, false);
    }
  }
  for (auto &g : merged) {
    std::sort(g.second.begin(), g.second.end());
    emit(g.second, true, g.first[3] != 0);
  }
  // hi/lo are indexed together; BIGCOEF classes index coef_fr instead
  if (out.lo.empty()) { out.lo.push_back(0); out.hi.push_back(0); }
  if (out.coef_fr.empty()) out.coef_fr.push_back(fr_zero());
}

// ---- fp_compile: rows -> the compiled program of kernels_r1cs_fast.cuh ---------------------------------------------------
struct fastprog_host {
  std::vector<uint32_t> bool_mask, bool_row, xor_ids, row_ids;
  std::vector<fp_xor> xors;
  std::vector<fp_tile> tiles, vtiles;    // vtiles: groups of 32 virtual-bit definitions (see fp_find_virtuals)
  std::vector<fp_item> items;
  uint32_t n_rows = 0, n_fast_tiles = 0, n_virtual = 0;
};

namespace r1cs_load_detail {
static int bitlen128(__int128 x) {
  unsigned __int128 m = x < 0 ? (unsigned __int128)(-x) : (unsigned __int128)x;
  int n = 0;
  while (m) { n++; m >>= 1; }
  return n;
}
// booleanity  (a x)(b x - b w0) = 0, either factor order
static bool is_bool_row(const row &R, uint32_t &x) {
  if (!R.part[2].empty()) return false;
  for (int sw = 0; sw < 2; sw++) {
    const std::vector<term> &P = R.part[sw], &Q = R.part[1 - sw];
    if (P.size() != 1 || Q.size() != 2 || !P[0].small || !Q[0].small || !Q[1].small) continue;
    const uint32_t w = P[0].wire;
    if (w == 0 || P[0].c == 0) continue;
    const term &qx = Q[0].wire == w ? Q[0] : Q[1], &q1 = Q[0].wire == w ? Q[1] : Q[0];
    if (qx.wire != w || q1.wire != 0 || qx.c == 0 || qx.c != -q1.c) continue;
    x = w;
    return true;
  }
  return false;
}
// XOR  (a x)(b y) = k x + k y - k o  with  a b = 2 k  (circom's 2 x y = x + y - out, any scaling or term order)
static bool is_xor_row(const row &R, uint32_t &x, uint32_t &y, uint32_t &o) {
  if (R.part[0].size() != 1 || R.part[1].size() != 1 || R.part[2].size() != 3) return false;
  const term &A = R.part[0][0], &B = R.part[1][0];
  if (!A.small || !B.small) return false;
  x = A.wire; y = B.wire;
  __int128 kx = 0, ky = 0, ko = 0;
  uint32_t seen = 0;
  for (const term &T : R.part[2]) {
    if (!T.small) return false;
    if (T.wire == x && !(seen & 1u)) { kx = T.c; seen |= 1u; }
    else if (T.wire == y && !(seen & 2u)) { ky = T.c; seen |= 2u; }
    else if (!(seen & 4u)) { ko = T.c; o = T.wire; seen |= 4u; }
    else return false;
  }
  const __int128 lim = (__int128)1 << 60;
  if (seen != 7u || x == y || o == x || o == y || x == 0 || y == 0 || o == 0) return false;
  if (A.c <= -lim || A.c >= lim || B.c <= -lim || B.c >= lim || kx == 0 || kx != ky || ko != -kx) return false;
  return A.c * B.c == 2 * kx;
}
// one linear combination -> items (scalars and runs of consecutive bit wires with doubling coefficients)
static bool compile_lc(std::vector<term> lc, std::vector<fp_item> &out) {
  std::sort(lc.begin(), lc.end(), [](const term &a, const term &b) { return a.wire < b.wire; });
  for (size_t j = 0; j < lc.size();) {
    if (!lc[j].small) return false;
    __int128 c = lc[j].c;
    if (c == 0) { j++; continue; }
    size_t len = 1;
    if (lc[j].wire != 0)
      while (j + len < lc.size() && len < 32 && lc[j + len].small && lc[j + len].wire == lc[j].wire + len && bitlen128(c) + (int)len < 120 &&
             lc[j + len].c == c * ((__int128)1 << len))
        len++;
    uint32_t shift = 0;
    while (bitlen128(c) > 61) {
      if (c & 1) return false;                             // does not fit  (62-bit mantissa) << shift
      c /= 2;
      shift++;
    }
    const int cbits = bitlen128(c) + (int)shift;
    if (cbits + 32 > 250 || shift > 200) return false;
    fp_item it;
    it.wire = lc[j].wire;
    it.meta = (uint32_t)(len >= 2 ? len : 0) | (shift << 8) | ((uint32_t)cbits << 16);
    it.coef = (long long)c;
    out.push_back(it);
    j += len;
  }
  return out.size() <= 255;
}
}  // namespace r1cs_load_detail

// ---- virtual bits ----------------------------------------------------------------------------------------------------
// circom's O2 pass removes a signal per linear constraint: a bit b_k of a Num2Bits / Bits34 decomposition disappears and
// every row that used it carries the linear combination  L = w - sum_{i != k} 2^i b_i  (= 2^k b_k) in its place -- a
// booleanity row becomes  (L)(L - u) = 0  with 32 terms a side, an XOR row drags one or two copies of L along.  Evaluated
// as they stand those rows are most of the stand-alone check's arithmetic for the O2 builds (and all of its 128-bit part).
// fp_find_virtuals recognises the rows  (L + a0)(s L + b0) = 0, s = +-1, one root zero  ->  L is 0 or u: V = L / u is a
// VIRTUAL BIT, evaluated once per witness into a bit slot past the witness's own (kernels_r1cs_fast.cuh, fp_eval_virtuals);
// fp_reduce_lc then rewrites  mu L + rest  ->  (mu u) V + rest  in every other row, after which those rows are booleanity /
// XOR / short rows like the rest of the system.  The rewriting is an identity whenever every L really is 0 or u; a witness
// for which one is not (a violated row) is evaluated by the program compiled WITHOUT virtual bits, so the verdict and the
// smallest violated row id never depend on this pass.
namespace r1cs_load_detail {
struct virt_def {
  std::vector<term> L;      // sorted by wire, no wire 0, powers of two divided out, first coefficient positive
  __int128 u;               // L is 0 or u
  uint32_t row;             // index (not id) of the defining row
  uint32_t anchor;          // position in L of a term with an odd coefficient
};
static int tz128(__int128 x) {
  unsigned __int128 m = x < 0 ? (unsigned __int128)(-x) : (unsigned __int128)x;
  int n = 0;
  while (m && !(m & 1)) { n++; m >>= 1; }
  return n;
}
static bool lc_small_sorted(const std::vector<term> &in, std::vector<term> &out, __int128 &c0) {
  out.clear();
  c0 = 0;
  for (const term &T : in) {
    if (!T.small) return false;
    if (T.wire == 0) c0 += T.c;
    else if (T.c != 0) out.push_back(T);
  }
  std::sort(out.begin(), out.end(), [](const term &a, const term &b) { return a.wire < b.wire; });
  for (size_t i = 1; i < out.size(); i++)
    if (out[i].wire == out[i - 1].wire) return false;        // a wire twice in one combination: not a form this pass meets
  return true;
}
static bool find_virtual(const row &R, uint32_t row_index, virt_def &V) {
  if (!R.part[2].empty() || R.part[0].size() < 3 || R.part[1].size() < 3) return false;
  std::vector<term> A, B;
  __int128 a0, b0;
  if (!lc_small_sorted(R.part[0], A, a0) || !lc_small_sorted(R.part[1], B, b0) || A.size() != B.size() || A.size() < 2) return false;
  const int sgn = B[0].c == A[0].c ? 1 : B[0].c == -A[0].c ? -1 : 0;
  if (!sgn) return false;
  for (size_t i = 0; i < A.size(); i++)
    if (A[i].wire != B[i].wire || B[i].c != sgn * A[i].c) return false;
  // (L + a0)(sgn L + b0) = 0:  L = -a0  or  L = -sgn b0
  __int128 u;
  if (a0 == 0 && b0 != 0) u = -sgn * b0;
  else if (b0 == 0 && a0 != 0) u = -a0;
  else return false;
  int g = 127;
  for (const term &T : A) g = std::min(g, tz128(T.c));
  if (tz128(u) < g) return false;                              // L is a multiple of 2^g and u is not: L = 0 always; leave the row alone
  const __int128 d = ((__int128)1 << g) * (A[0].c < 0 ? -1 : 1);
  V.L = A;
  for (term &T : V.L) T.c /= d;
  V.u = u / d;
  V.row = row_index;
  V.anchor = 0;
  while (V.anchor < V.L.size() && !(V.L[V.anchor].c & 1)) V.anchor++;
  return V.anchor < V.L.size();
}
// rewrite every  mu * L_j  inside `lc` as  (mu u_j) * V_j  (V_j = wire vbase + j); returns true when something changed
static bool reduce_lc(std::vector<term> &lc, const std::vector<virt_def> &virt, const std::multimap<uint32_t, uint32_t> &by_anchor, uint32_t vbase) {
  bool changed = false;
  for (bool again = true; again;) {
    again = false;
    std::map<uint32_t, __int128> m;
    for (const term &T : lc) {
      if (!T.small) return changed;
      m[T.wire] += T.c;
    }
    for (const auto &kv : m) {
      auto range = by_anchor.equal_range(kv.first);
      for (auto it = range.first; it != range.second && !again; ++it) {
        const virt_def &V = virt[it->second];
        const __int128 ca = V.L[V.anchor].c, mu = kv.second / ca;
        if (mu == 0 || mu * ca != kv.second || bitlen128(mu) + bitlen128(V.u) > 118) continue;
        bool all = true;
        for (const term &T : V.L) {
          auto f = m.find(T.wire);
          all = all && f != m.end() && bitlen128(mu) + bitlen128(T.c) < 120 && f->second == mu * T.c;
        }
        if (!all) continue;
        for (const term &T : V.L) m.erase(T.wire);
        m[vbase + it->second] += mu * V.u;
        lc.clear();
        for (const auto &e : m)
          if (e.second != 0) {
            term T;
            T.wire = e.first; T.small = true; T.c = e.second; T.f = fr_zero();
            lc.push_back(T);
          }
        changed = again = true;
      }
      if (again) break;
    }
  }
  return changed;
}
}  // namespace r1cs_load_detail

// Compiles what it can; taken[i] tells which rows are now covered by the program (the others go to r1cs_group).
// wide_hint (may be NULL): per wire, 1 where the circuit's slot kinds say the value is a field element (IsZero's inverse):
// a tile with such a scalar is not marked FP_TILE_FAST -- at run time its 64-bit
// pass would find the value out of bounds and the tile would be evaluated a second time.  A hint only: the verdict of a tile
// never depends on its flag.
static void fp_compile(const std::vector<r1cs_load_detail::row> &rows, uint32_t ws, fastprog_host &fp, std::vector<char> &taken,
                       bool with_virtuals = false, const std::vector<uint8_t> *wide_hint = nullptr) {
  using namespace r1cs_load_detail;
  const uint32_t words = (ws + 31u) >> 5, vbase = words * 32u;
  taken.assign(rows.size(), 0);
  auto compile_row = [](const row &R, std::vector<fp_item> &it, uint32_t n[3]) {
    it.clear();
    for (int part = 0; part < 3; part++) {
      std::vector<fp_item> tmp;
      if (!compile_lc(R.part[part], tmp)) return false;
      it.insert(it.end(), tmp.begin(), tmp.end());
      n[part] = (uint32_t)tmp.size();
    }
    return true;
  };
  // virtual bits: the rows that define them, and their compiled linear combinations
  std::vector<virt_def> virt;
  std::vector<std::vector<fp_item>> virt_items;
  std::vector<int> def_of(rows.size(), -1);
  std::multimap<uint32_t, uint32_t> by_anchor;
  if (with_virtuals)
    for (size_t i = 0; i < rows.size(); i++) {
      virt_def V;
      std::vector<fp_item> it, li;
      uint32_t n[3];
      uint32_t x;
      if (is_bool_row(rows[i], x) || !find_virtual(rows[i], (uint32_t)i, V)) continue;
      if (!compile_row(rows[i], it, n) || !compile_lc(V.L, li) || li.empty() || li.size() > 16 || bitlen128(V.u) > 118) continue;
      if (bitlen128(V.u) - tz128(V.u) > 61) continue;            // the unit is stored as (62-bit mantissa) << shift
      def_of[i] = (int)virt.size();
      by_anchor.insert({V.L[V.anchor].wire, (uint32_t)virt.size()});
      virt.push_back(std::move(V));
      virt_items.push_back(std::move(li));
    }
  const uint32_t nvw = ((uint32_t)virt.size() + 31u) >> 5;
  fp.n_virtual = (uint32_t)virt.size();
  fp.bool_mask.assign(words + nvw + 1, 0u);
  fp.bool_row.assign((size_t)(words + nvw) * 32, 0xFFFFFFFFu);
  struct xr { uint32_t x, y, o, id; };
  std::vector<xr> xs;
  struct gen { std::vector<fp_item> it; uint32_t n[3]; uint32_t id; };
  std::vector<gen> gens;
  for (size_t i = 0; i < rows.size(); i++) {
    if (def_of[i] >= 0) { taken[i] = 1; continue; }            // holds by construction once the virtual bit is valid
    const row *Rp = &rows[i];
    row reduced;
    if (!virt.empty() && (rows[i].part[0].size() > 2 || rows[i].part[1].size() > 2 || rows[i].part[2].size() > 2)) {
      reduced = rows[i];
      bool changed = false;
      for (int part = 0; part < 3; part++) changed = reduce_lc(reduced.part[part], virt, by_anchor, vbase) || changed;
      if (changed) {
        std::vector<fp_item> it;
        uint32_t n[3];
        if (!compile_row(rows[i], it, n)) continue;            // the fallback program must cover the same rows: leave it to the general evaluator
        if (reduced.part[0].empty() || reduced.part[1].empty()) { reduced.part[0].clear(); reduced.part[1].clear(); }
        Rp = &reduced;
      }
    }
    const row &R = *Rp;
    uint32_t x = 0, y = 0, o = 0;
    if (is_bool_row(R, x)) {
      fp.bool_mask[x >> 5] |= 1u << (x & 31u);
      fp.bool_row[x] = std::min(fp.bool_row[x], R.id);
      taken[i] = 1;
    } else if (is_xor_row(R, x, y, o)) {
      xs.push_back(xr{x, y, o, R.id});
      taken[i] = 1;
    } else {
      gen g;
      g.id = R.id;
      bool ok = compile_row(R, g.it, g.n);
      if (!ok && Rp == &reduced) ok = compile_row(rows[i], g.it, g.n);
      if (ok) { gens.push_back(std::move(g)); taken[i] = 1; }
    }
  }
  // virtual-bit definitions -> groups of 32 (lane = definition): its items, item-major, padded to the group's longest, then
  // one more item holding the unit u as coef << shift
  for (size_t j = 0; j < virt.size(); j += 32) {
    const size_t cnt = std::min<size_t>(32, virt.size() - j);
    size_t ni = 0;
    for (size_t l = 0; l < cnt; l++) ni = std::max(ni, virt_items[j + l].size());
    fp_tile t;
    t.item_off = (uint32_t)fp.items.size();
    t.row_off = (uint32_t)fp.row_ids.size();
    t.nA = (uint16_t)ni; t.nB = 0; t.nC = 0; t.rows = (uint16_t)cnt;
    const fp_item pad = {0u, 1u << 24, 0ll};                 // wire 0 times 0 (its value bound: one bit)
    for (size_t k = 0; k < ni; k++)
      for (size_t l = 0; l < 32; l++) fp.items.push_back(l < cnt && k < virt_items[j + l].size() ? virt_items[j + l][k] : pad);
    for (size_t l = 0; l < 32; l++) {
      fp_item ui = pad;
      if (l < cnt) {
        __int128 u = virt[j + l].u;                            // (62-bit mantissa) << shift, exactly: checked when the row was picked
        uint32_t shift = 0;
        while (bitlen128(u) > 61) { u /= 2; shift++; }
        ui.wire = 0; ui.meta = shift << 8; ui.coef = (long long)u;
      }
      fp.items.push_back(ui);
    }
    for (size_t l = 0; l < 32; l++) fp.row_ids.push_back(l < cnt ? rows[virt[j + l].row].id : 0xFFFFFFFFu);
    fp.vtiles.push_back(t);
  }
  // XOR rows -> runs
  std::sort(xs.begin(), xs.end(), [](const xr &a, const xr &b) { return a.x != b.x ? a.x < b.x : a.id < b.id; });
  for (size_t j = 0; j < xs.size();) {
    size_t len = 1;
    while (j + len < xs.size() && len < 32 && xs[j + len].x == xs[j].x + len && xs[j + len].y == xs[j].y + len && xs[j + len].o == xs[j].o + len) len++;
    fp_xor e;
    e.x = xs[j].x; e.y = xs[j].y; e.o = xs[j].o;
    e.len_id = (uint32_t)len | ((uint32_t)fp.xor_ids.size() << 6);
    for (size_t k = 0; k < len; k++) fp.xor_ids.push_back(xs[j + k].id);
    fp.xors.push_back(e);
    j += len;
  }
  // general rows -> tiles of 32 rows with identical item counts, items item-major
  std::stable_sort(gens.begin(), gens.end(), [](const gen &a, const gen &b) {
    if (a.n[0] != b.n[0]) return a.n[0] < b.n[0];
    if (a.n[1] != b.n[1]) return a.n[1] < b.n[1];
    return a.n[2] < b.n[2];
  });
  struct tile_plan { std::vector<size_t> rows; uint32_t n[3]; bool fast; uint32_t cost; };
  std::vector<tile_plan> plan, small;
  for (size_t j = 0; j < gens.size();) {
    size_t cnt = 1;
    while (j + cnt < gens.size() && cnt < 32 && gens[j + cnt].n[0] == gens[j].n[0] && gens[j + cnt].n[1] == gens[j].n[1] && gens[j + cnt].n[2] == gens[j].n[2]) cnt++;
    tile_plan tp;
    for (size_t l = 0; l < cnt; l++) tp.rows.push_back(j + l);
    for (int q = 0; q < 3; q++) tp.n[q] = gens[j].n[q];
    (cnt <= 4 ? small : plan).push_back(std::move(tp));
    j += cnt;
  }
  // shapes with a handful of rows (the one-off rows of the nova step logic) would cost a warp pass each: they share
  // tiles, every row padded with zero items to the tile's longest A, B and C
  std::stable_sort(small.begin(), small.end(), [](const tile_plan &a, const tile_plan &b) { return a.n[0] + a.n[1] + a.n[2] > b.n[0] + b.n[1] + b.n[2]; });
  for (size_t j = 0; j < small.size();) {
    tile_plan tp = small[j++];
    while (j < small.size() && tp.rows.size() + small[j].rows.size() <= 32) {
      for (int q = 0; q < 3; q++) tp.n[q] = std::max(tp.n[q], small[j].n[q]);
      tp.rows.insert(tp.rows.end(), small[j].rows.begin(), small[j].rows.end());
      j++;
    }
    plan.push_back(std::move(tp));
  }
  // what the 64-bit path may assume of a scalar (and checks at run time): a wire with a booleanity row, a virtual bit and
  // wire 0 stay below 2^1, anything else below 2^FP_FAST_VBITS; kept in bits 24..29 of the item's meta word
  for (gen &g : gens)
    for (fp_item &it : g.it)
      if ((it.meta & 63u) == 0) {
        const bool bit = it.wire == 0 || it.wire >= vbase || ((fp.bool_mask[it.wire >> 5] >> (it.wire & 31u)) & 1u);
        it.meta |= (bit ? 1u : (uint32_t)FP_FAST_VBITS) << 24;
      }
  for (tile_plan &tp : plan) {
    // FP_TILE_FAST: 64-bit sums are exact when every scalar is below its bound and every run lies over bits
    bool fast = tp.n[0] <= 16 && tp.n[1] <= 16 && tp.n[2] <= 16, scalar_product = false;
    for (size_t r : tp.rows) {
      for (const fp_item &it : gens[r].it) {
        const uint32_t len = it.meta & 63u, cbits = (it.meta >> 16) & 255u;
        fast = fast && cbits + (len ? len : (it.meta >> 24) & 63u) <= 57;
        if (wide_hint && len == 0 && it.coef != 0 && it.wire < wide_hint->size() && (*wide_hint)[it.wire]) fast = false;
      }
      bool sp = gens[r].n[0] && gens[r].n[1];
      for (uint32_t k = 0; k < gens[r].n[0] + gens[r].n[1]; k++) sp = sp && (gens[r].it[k].meta & 63u) == 0;
      scalar_product = scalar_product || sp;
    }
    // cost estimate for the hand-out order: items, dearer on the bounds-tracking path; products of plain wires are where
    // field-valued operands turn up (IsZero's in * inv), i.e. rows that may need the Fr evaluator
    tp.fast = fast;
    tp.cost = (tp.n[0] + tp.n[1] + tp.n[2]) * (fast ? 2u : 5u) + (scalar_product ? 64u : 0u);
  }
  std::stable_sort(plan.begin(), plan.end(), [](const tile_plan &a, const tile_plan &b) { return a.cost > b.cost; });
  for (const tile_plan &tp : plan) {
    const size_t cnt = tp.rows.size();
    fp_tile t;
    t.item_off = (uint32_t)fp.items.size();
    t.row_off = (uint32_t)fp.row_ids.size();
    t.nA = (uint16_t)tp.n[0]; t.nB = (uint16_t)tp.n[1]; t.nC = (uint16_t)tp.n[2];
    t.rows = (uint16_t)(cnt | (tp.fast ? FP_TILE_FAST : 0u));
    const fp_item pad = {0u, 1u << 24, 0ll};                 // wire 0 times 0 (its value bound: one bit)
    for (int q = 0; q < 3; q++)
      for (uint32_t k = 0; k < tp.n[q]; k++)
        for (uint32_t l = 0; l < 32; l++) {
          fp_item it = pad;
          if (l < cnt) {
            const gen &g = gens[tp.rows[l]];
            const uint32_t base = q == 0 ? 0u : q == 1 ? g.n[0] : g.n[0] + g.n[1];
            if (k < g.n[q]) it = g.it[base + k];
          }
          fp.items.push_back(it);
        }
    for (uint32_t l = 0; l < 32; l++) fp.row_ids.push_back(l < cnt ? gens[tp.rows[l]].id : 0xFFFFFFFFu);
    fp.tiles.push_back(t);
    fp.n_fast_tiles += tp.fast ? 1u : 0u;
  }
  for (char c : taken) fp.n_rows += c ? 1u : 0u;
}

// Prepare a slot-space set for the general evaluator (r1cs_rows.cuh).  Every class is cut into row blocks: <= 32
// consecutive rows in which each term column is an arithmetic progression; a class whose blocks would average fewer
// than 8 rows keeps its [term][row] matrix instead (flag MATRIX).  In: cls[k].term_off = start of the class's matrix in
// `terms`.  Out: `blocks` = the new term store (per block: first row, rows, then {first wire, wire step} per term; or
// the matrix), cls[k].term_off = start of the class's data in it (even: the pairs are 8-byte aligned), nblk[k] = number
// of blocks (0 for MATRIX classes); flag COEF64 where every coefficient of the class fits int64.
static void stg_blockify(std::vector<r1cs_class_dev> &cls, const std::vector<uint32_t> &terms, std::vector<uint32_t> &blocks,
                         std::vector<uint32_t> &nblk, const std::vector<int64_t> &lo, const std::vector<int64_t> &hi) {
  blocks.clear();
  nblk.assign(cls.size(), 0);
  for (size_t k = 0; k < cls.size(); k++) {
    r1cs_class_dev &c = cls[k];
    const uint32_t nt = (uint32_t)c.nA + c.nB + c.nC;
    const uint32_t *m = terms.data() + c.term_off;
    if (!(c.flags & R1CS_FLAG_BIGCOEF)) {
      const size_t ncoef = (c.flags & R1CS_FLAG_ROWCOEF) ? (size_t)nt * c.count : nt;
      bool fits = true, fits56 = true;
      for (size_t i = 0; i < ncoef && fits; i++) {
        const int64_t l = lo[c.coef_off + i];
        fits = hi[c.coef_off + i] == (l < 0 ? -1 : 0);
        fits56 = fits56 && fits && l > -(1ll << 56) && l < (1ll << 56);
      }
      if (fits) c.flags |= R1CS_FLAG_COEF64;
      // the 64-bit evaluator (r1cs_rows.cuh, STG_FAST_*): <= 64 terms of < 2^56 each stay below 2^62
      if (fits && fits56 && c.nA <= 64 && c.nB <= 64 && c.nC <= 64) c.flags |= R1CS_FLAG_FAST64;
      // booleanity rows  (a x) * (b x - b w0) = 0  with w0 = wire 0, the constant 1:  "x is 0 or 1"
      if (fits && c.nA == 1 && c.nB == 2 && c.nC == 0 && c.count) {
        bool boolrow = true;
        for (uint32_t r = 0; r < c.count && boolrow; r++) {
          auto co = [&](uint32_t t) { return lo[c.coef_off + ((c.flags & R1CS_FLAG_ROWCOEF) ? (size_t)t * c.count + r : t)]; };
          const uint32_t x = m[r], w1 = m[(size_t)c.count + r], w2 = m[(size_t)2 * c.count + r];
          const int64_t a = co(0), b1 = co(1), b2 = co(2);
          const bool fwd = w1 == x && w2 == 0, rev = w2 == x && w1 == 0;        // which B term is x, which is the constant
          boolrow = x != 0 && a != 0 && b1 != 0 && b1 == -b2 && b1 != INT64_MIN && (fwd || rev);
        }
        if (boolrow) c.flags |= R1CS_FLAG_BOOLROW;
      }
      // XOR rows  (a x)(b y) = k x + k y - k o  with a b = 2 k  (circom's  2 x y = x + y - out  for bit operands, any scaling
      // or term order): over bits it says o = x xor y.  The device finds o as the C wire that is neither x nor y.
      if (fits && c.nA == 1 && c.nB == 1 && c.nC == 3 && c.count) {
        bool xorrow = true;
        for (uint32_t r = 0; r < c.count && xorrow; r++) {
          auto co = [&](uint32_t t) { return lo[c.coef_off + ((c.flags & R1CS_FLAG_ROWCOEF) ? (size_t)t * c.count + r : t)]; };
          const uint32_t x = m[r], y = m[(size_t)c.count + r];
          const int64_t a = co(0), b = co(1);
          int64_t kx = 0, ky = 0, ko = 0;
          uint32_t seen = 0, o = 0;
          for (uint32_t t = 2; t < 5; t++) {
            const uint32_t wq = m[(size_t)t * c.count + r];
            if (wq == x && !(seen & 1u)) { kx = co(t); seen |= 1u; }
            else if (wq == y && !(seen & 2u)) { ky = co(t); seen |= 2u; }
            else if (!(seen & 4u)) { ko = co(t); o = wq; seen |= 4u; }
            else seen |= 8u;
          }
          const bool small = a > -(1ll << 30) && a < (1ll << 30) && b > -(1ll << 30) && b < (1ll << 30);
          xorrow = seen == 7u && x != y && o != x && o != y && x != 0 && y != 0 && o != 0 && small && kx != 0 && kx == ky &&
                   ko == -kx && a * b == 2 * kx;
        }
        if (xorrow) c.flags |= R1CS_FLAG_XORROW;
      }
    }
    if (blocks.size() & 1) blocks.push_back(0);
    std::vector<uint32_t> mine;
    uint32_t r = 0, nb = 0;
    while (r < c.count) {
      uint32_t len = 1;
      if (r + 1 < c.count) {
        len = 2;
        auto step = [&](uint32_t t) { return m[(size_t)t * c.count + r + 1] - m[(size_t)t * c.count + r]; };
        while (r + len < c.count && len < 32) {
          bool same = true;
          for (uint32_t t = 0; t < nt && same; t++)
            same = m[(size_t)t * c.count + r + len] - m[(size_t)t * c.count + r + len - 1] == step(t);
          if (!same) break;
          len++;
        }
      }
      mine.push_back(r);
      mine.push_back(len);
      for (uint32_t t = 0; t < nt; t++) {
        mine.push_back(m[(size_t)t * c.count + r]);
        mine.push_back(len > 1 ? m[(size_t)t * c.count + r + 1] - m[(size_t)t * c.count + r] : 0u);
      }
      nb++;
      r += len;
    }
    c.term_off = (uint32_t)blocks.size();
    if (c.count && (uint64_t)nb * 8 > c.count) {            // short blocks: the matrix is the better form
      c.flags |= R1CS_FLAG_MATRIX;
      blocks.insert(blocks.end(), m, m + (size_t)nt * c.count);
    } else {
      nblk[k] = nb;
      blocks.insert(blocks.end(), mine.begin(), mine.end());
    }
  }
  if (blocks.empty()) blocks.push_back(0);
}
