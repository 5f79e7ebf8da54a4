// vnnlib -> batched (box, S) safety queries (host only).
//
// Reads the "simple" vnnlib subset the reference reads with read_vnnlib_simple
// (exts/vnnlib_parser.jl:99-216) and flattens it the way loadVnnlibCnf does
// (experiments/vnnlib_utils.jl:18-56): the file states NOT phi in disjunctive normal form,
//     OR_{box} OR_{(A, b)} (x in box  AND  A y <= b),
// and the property is verified as the conjunction, over every (box, (A, b)), of the disjunctive clause
//     OR_i (x in box  AND  A_i y >= b_i + eps),       eps = 1e-4,
// one SafetyQuery per row i with S = hplaneS(-A_i, -b_i - eps) (src/Utils/qc.jl:27-37).  The output arrays are
// what nnsdp_assemble_blocks takes (x1min / x1max with stride n_in, out_S with stride sdim^2), plus the clause
// index of every query, so an ACAS property is one batch.
//
// Statement handling follows the reference reader: ';' starts a comment, a statement may span lines (balanced
// parentheses), `(declare-const X_i Real)` is skipped, `(assert (<=|>= a b))` applies to every alternative, and
// `(assert (or (and c...)(and c...)...))` multiplies the alternatives.  X_i constraints tighten the box
// (<= : upper, >= : lower); Y constraints become a row of A y <= b (>= swaps its operands first).  Alternatives
// with the same box are merged; the reference iterates a Dict there (hash order), here boxes keep the order of
// first appearance.  Anything else is rejected with NNSDP_ERR_ARG instead of being fed to the DNF branch.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "internal.h"

using namespace nnsdp;

namespace {

struct Alt {                       // one alternative of the DNF: a box and rows of A y <= b
  std::vector<double> lo, hi;
  std::vector<std::vector<double>> rows;
  std::vector<double> rhs;
};

std::string strip(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && isspace((unsigned char)s[i])) ++i;
    size_t j = i;
    while (j < s.size() && !isspace((unsigned char)s[j])) ++j;
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j;
  }
  return out;
}

void replace_all(std::string* s, const std::string& from, const std::string& to) {
  size_t pos = 0;
  while ((pos = s->find(from, pos)) != std::string::npos) {
    s->replace(pos, from.size(), to);
    pos += to.size();
  }
}

// vnnlib_parser.jl:3-44
int32_t read_statements(const char* path, std::vector<std::string>* out) {
  FILE* f = fopen(path, "r");
  NN_CHECK(f != nullptr, NNSDP_ERR_ARG, "cannot open %s", path);
  std::string text;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  long open = 0;
  std::string cur;
  size_t pos = 0;
  bool any = false;
  while (pos <= text.size()) {
    size_t nl = text.find('\n', pos);
    if (nl == std::string::npos) nl = text.size();
    std::string line = strip(text.substr(pos, nl - pos));
    pos = nl + 1;
    any = true;
    const size_t sc = line.find(';');
    if (sc != std::string::npos) line = strip(line.substr(0, sc));
    if (line.empty()) continue;
    for (char c : line) open += (c == '(') - (c == ')');
    NN_CHECK(open >= 0, NNSDP_ERR_ASSERT, "%s: unbalanced ')' (vnnlib_parser.jl:23)", path);
    if (!cur.empty()) cur += " ";
    cur += line;
    if (open == 0) {
      out->push_back(cur);
      cur.clear();
    }
  }
  NN_CHECK(any, NNSDP_ERR_ASSERT, "%s is empty (vnnlib_parser.jl:6)", path);
  if (!cur.empty()) out->push_back(cur);
  for (std::string& s : *out) {
    const std::vector<std::string> tk = split_ws(s);
    s.clear();
    for (size_t i = 0; i < tk.size(); ++i) s += (i ? " " : "") + tk[i];
    replace_all(&s, "( ", "(");
    replace_all(&s, ") ", ")");
  }
  return NNSDP_OK;
}

bool parse_index(const std::string& tok, long* idx) {   // "X_12" / "Y_3" -> 12 / 3
  if (tok.size() < 3) return false;
  char* end = nullptr;
  *idx = strtol(tok.c_str() + 2, &end, 10);
  return end != tok.c_str() + 2 && *end == '\0';
}
bool parse_real(const std::string& tok, double* v) {
  char* end = nullptr;
  *v = strtod(tok.c_str(), &end);
  return end != tok.c_str() && *end == '\0';
}
bool starts(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }

// update_rv_tuple!, vnnlib_parser.jl:46-92
int32_t apply(Alt* a, const std::string& op, std::string first, std::string second, int n_in, int n_out) {
  NN_CHECK(op == "<=" || op == ">=", NNSDP_ERR_ARG, "vnnlib: unsupported comparison '%s'", op.c_str());
  long i1 = 0, i2 = 0;
  double v = 0.0;
  if (starts(first, "X_")) {
    NN_CHECK(parse_index(first, &i1), NNSDP_ERR_ARG, "vnnlib: bad variable '%s'", first.c_str());
    NN_CHECK(!starts(second, "X") && !starts(second, "Y"), NNSDP_ERR_ASSERT,
             "vnnlib: input constraint against a variable (vnnlib_parser.jl:51)");
    NN_CHECK(0 <= i1 && i1 < n_in, NNSDP_ERR_ASSERT, "vnnlib: X_%ld out of range (vnnlib_parser.jl:52)", i1);
    NN_CHECK(parse_real(second, &v), NNSDP_ERR_ARG, "vnnlib: bad number '%s'", second.c_str());
    if (op == "<=") a->hi[i1] = fmin(v, a->hi[i1]);
    else a->lo[i1] = fmax(v, a->lo[i1]);
    NN_CHECK(a->lo[i1] <= a->hi[i1], NNSDP_ERR_ASSERT, "vnnlib: empty interval for X_%ld (vnnlib_parser.jl:60)", i1);
    return NNSDP_OK;
  }
  if (op == ">=") first.swap(second);
  std::vector<double> row(n_out, 0.0);
  double rhs = 0.0;
  const bool y1 = starts(first, "Y_"), y2 = starts(second, "Y_");
  if (y1) {
    NN_CHECK(parse_index(first, &i1) && 0 <= i1 && i1 < n_out, NNSDP_ERR_ARG, "vnnlib: bad output '%s'", first.c_str());
  }
  if (y2) {
    NN_CHECK(parse_index(second, &i2) && 0 <= i2 && i2 < n_out, NNSDP_ERR_ARG, "vnnlib: bad output '%s'", second.c_str());
  }
  if (y1 && y2) {
    row[i1] = 1.0;
    row[i2] = -1.0;                // Y_a <= Y_a leaves -1, like the reference's two assignments
  } else if (y1) {
    NN_CHECK(parse_real(second, &rhs), NNSDP_ERR_ARG, "vnnlib: bad number '%s'", second.c_str());
    row[i1] = 1.0;
  } else {
    NN_CHECK(y2, NNSDP_ERR_ASSERT, "vnnlib: comparison without an output variable (vnnlib_parser.jl:83)");
    NN_CHECK(parse_real(first, &v), NNSDP_ERR_ARG, "vnnlib: bad number '%s'", first.c_str());
    row[i2] = -1.0;
    rhs = -1.0 * v;
  }
  a->rows.push_back(row);
  a->rhs.push_back(rhs);
  return NNSDP_OK;
}

// "(<= a b)" -> (op, a, b); false unless the whole string is one comparison
bool parse_comparison(const std::string& s, std::string* op, std::string* a, std::string* b) {
  if (s.size() < 2 || s.front() != '(' || s.back() != ')') return false;
  const std::string in = s.substr(1, s.size() - 2);
  if (in.find('(') != std::string::npos || in.find(')') != std::string::npos) return false;
  const std::vector<std::string> tk = split_ws(in);
  if (tk.size() != 3) return false;
  *op = tk[0], *a = tk[1], *b = tk[2];
  return *op == "<=" || *op == ">=";
}

int32_t parse_file(const char* path, int n_in, int n_out, std::vector<Alt>* boxes_out) {
  std::vector<std::string> st;
  NN_TRY(read_statements(path, &st));
  std::vector<Alt> alts(1);
  alts[0].lo.assign(n_in, -INFINITY);
  alts[0].hi.assign(n_in, INFINITY);
  for (const std::string& s : st) {
    if (starts(s, "(declare-const ")) continue;
    NN_CHECK(starts(s, "(assert ") && s.back() == ')', NNSDP_ERR_ARG, "vnnlib: unsupported statement '%.60s'", s.c_str());
    const std::string body = s.substr(8, s.size() - 9);   // inside (assert ...)
    std::string op, a, b;
    if (parse_comparison(body, &op, &a, &b)) {             // regex_simple_assert
      for (Alt& alt : alts) NN_TRY(apply(&alt, op, a, b, n_in, n_out));
      continue;
    }
    NN_CHECK(starts(body, "(or (and "), NNSDP_ERR_ARG,
             "vnnlib: only (assert (<=|>= a b)) and (assert (or (and ...)...)) are supported: '%.60s'", s.c_str());
    // tokens without parentheses, after 'or'; conjuncts are separated by the token 'and'
    std::string flat = body;
    for (char& c : flat)
      if (c == '(' || c == ')') c = ' ';
    std::vector<std::string> tk = split_ws(flat);           // or and op a b op a b and ...
    std::vector<std::vector<std::string>> conj;
    for (size_t i = 1; i < tk.size(); ++i) {
      if (tk[i] == "and") conj.emplace_back();
      else {
        NN_CHECK(!conj.empty(), NNSDP_ERR_ARG, "vnnlib: malformed disjunction '%.60s'", s.c_str());
        conj.back().push_back(tk[i]);
      }
    }
    std::vector<Alt> next;
    for (const Alt& alt : alts)
      for (const std::vector<std::string>& c : conj) {
        Alt cp = alt;
        for (size_t k = 0; k < c.size() / 3; ++k)   // count = length / 3 (vnnlib_parser.jl:158)
          NN_TRY(apply(&cp, c[3 * k], c[3 * k + 1], c[3 * k + 2], n_in, n_out));
        next.push_back(std::move(cp));
      }
    alts.swap(next);
  }
  // merge alternatives with the same box (first appearance order)
  std::vector<Alt>& boxes = *boxes_out;
  std::vector<std::vector<size_t>> members;
  for (size_t i = 0; i < alts.size(); ++i) {
    size_t j = 0;
    for (; j < boxes.size(); ++j)
      if (boxes[j].lo == alts[i].lo && boxes[j].hi == alts[i].hi) break;
    if (j == boxes.size()) {
      Alt bx;
      bx.lo = alts[i].lo, bx.hi = alts[i].hi;
      boxes.push_back(bx);
      members.emplace_back();
    }
    members[j].push_back(i);
  }
  // the specs (mat, rhs) of a box, one per alternative, back to back; an empty row ends a spec
  for (size_t j = 0; j < boxes.size(); ++j) {
    for (int d = 0; d < n_in; ++d)
      NN_CHECK(isfinite(boxes[j].lo[d]) && isfinite(boxes[j].hi[d]), NNSDP_ERR_ASSERT,
               "vnnlib: X_%d is unbounded (vnnlib_parser.jl:200)", d);
    for (size_t i : members[j]) {
      NN_CHECK(!alts[i].rows.empty(), NNSDP_ERR_ASSERT, "vnnlib: an alternative has no output constraint");
      for (size_t r = 0; r < alts[i].rows.size(); ++r) {
        boxes[j].rows.push_back(alts[i].rows[r]);
        boxes[j].rhs.push_back(alts[i].rhs[r]);
      }
      boxes[j].rows.emplace_back();          // empty row = end of this spec
      boxes[j].rhs.push_back(0.0);
    }
  }
  return NNSDP_OK;
}

}  // namespace

extern "C" int32_t nnsdp_vnnlib_read(const char* path, int64_t n_in, int64_t n_out, int64_t max_queries,
                                     int64_t* nqueries, int64_t* nclauses, double* x1min, double* x1max,
                                     double* out_S, int64_t* clause_of) {
  NN_CHECK(path && nqueries && nclauses, NNSDP_ERR_ARG, "NULL argument");
  NN_CHECK(n_in >= 1 && n_out >= 1, NNSDP_ERR_ARG, "n_in and n_out must be >= 1");
  std::vector<Alt> boxes;
  NN_TRY(parse_file(path, (int)n_in, (int)n_out, &boxes));
  const int64_t sdim = n_in + n_out + 1;
  const double eps = 1e-4;   // vnnlib_utils.jl:41
  int64_t nq = 0, nc = 0;
  const bool fill = (x1min != nullptr);
  if (fill) NN_CHECK(x1max && out_S && clause_of, NNSDP_ERR_ARG, "NULL output array");
  for (const Alt& bx : boxes) {
    for (size_t r = 0; r < bx.rows.size(); ++r) {
      if (bx.rows[r].empty()) {           // end of one (mat, rhs) spec = one disjunctive clause
        nc += 1;
        continue;
      }
      if (fill) {
        NN_CHECK(nq < max_queries, NNSDP_ERR_ARG, "max_queries = %lld is too small", (long long)max_queries);
        for (int64_t d = 0; d < n_in; ++d) {
          x1min[nq * n_in + d] = bx.lo[d];
          x1max[nq * n_in + d] = bx.hi[d];
        }
        // S = hplaneS(normal = -A_i, h = -b_i - eps): S23 = normal, S33 = -2 h  (src/Utils/qc.jl:27-37)
        double* S = out_S + nq * sdim * sdim;
        for (int64_t i = 0; i < sdim * sdim; ++i) S[i] = 0.0;
        for (int64_t o = 0; o < n_out; ++o) {
          const double nrm = -bx.rows[r][o];
          S[(n_in + o) + (sdim - 1) * sdim] = nrm;
          S[(sdim - 1) + (n_in + o) * sdim] = nrm;
        }
        S[(sdim - 1) + (sdim - 1) * sdim] = -2.0 * (-bx.rhs[r] - eps);
        clause_of[nq] = nc;
      }
      nq += 1;
    }
  }
  *nqueries = nq;
  *nclauses = nc;
  return NNSDP_OK;
}
