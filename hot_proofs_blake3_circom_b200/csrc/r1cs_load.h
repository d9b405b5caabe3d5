// r1cs_load.h -- host side of b3w_r1cs_load: parse an iden3 `.r1cs` file (binary format v1: magic "r1cs", sections
// 1 header / 2 constraints / 3 wire->label map; what circom writes next to the wasm and what snarkjs, circom_tester and
// circom-scotia read -- rust_fold/src/blake3_circuit.rs:71-81) and group its rows into the shape classes the device
// evaluator works on (r1cs.cuh): rows with the same (nA, nB, nC) and the same small coefficients share a coefficient
// vector; small groups are merged into per-row-coefficient classes; rows whose coefficients do not fit 56 bits (e.g.
// 2^-31 mod p after circom's O2 substitution) keep full field-element coefficients (BIGCOEF) and are evaluated in Fr.
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
struct term { uint32_t wire; bool small; __int128 c; fr_t f; };
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

// returns 0 or a B3W_ERR_* code with the text in `err`
static int r1cs_parse(const uint8_t *data, size_t len, const uint8_t prime[32], uint32_t ws, r1cs_host_set &out, std::string &err) {
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
  std::vector<row> rows(m);
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
        // small signed form: c < 2^56 or p - c < 2^56
        fr_t neg;
        fr_raw_sub(neg, P, T.f);
        auto fits = [](const fr_t &x) { return (x.l[2] | x.l[3] | x.l[4] | x.l[5] | x.l[6] | x.l[7]) == 0 && (x.l[1] >> 24) == 0; };
        if (fits(T.f)) { T.small = true; T.c = (__int128)(((uint64_t)T.f.l[1] << 32) | T.f.l[0]); }
        else if (fits(neg)) { T.small = true; T.c = -(__int128)(((uint64_t)neg.l[1] << 32) | neg.l[0]); }
        else { T.small = false; T.c = 0; R.big = true; }
      }
      if (k > 64) R.big = true;                            // keeps the 128-bit accumulators of the integer path exact
    }
    if (R.part[0].empty() || R.part[1].empty()) { R.part[0].clear(); R.part[1].clear(); }      // 0 * B = C  <=>  C = 0
    for (int part = 0; part < 3; part++)
      std::sort(R.part[part].begin(), R.part[part].end(), [&](const term &a, const term &b) {
        if (!R.big && a.c != b.c) return a.c < b.c;
        return a.wire < b.wire;
      });
  }
  // group
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
  return B3W_OK;
}

// Prepare a slot-space set for the staged checker (kernels_r1cs_staged.cuh).  Every class is cut into row blocks: <= 32
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
      // the 64-bit evaluator (kernels_r1cs_staged.cuh, STG_FAST_*): <= 64 terms of < 2^56 each stay below 2^62
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
