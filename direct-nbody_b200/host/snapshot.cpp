// snapshot.cpp — host mirror of the reference's text snapshot format: the `--- !Particle`
// records written by Advancer.A.print (advancer.ml:88-96) and Leapfrog.print
// (leapfrog.ml:44-53) and read back by A.read (advancer.ml:118-124) and Leapfrog.read
// (leapfrog.ml:55-62).  OCaml's Printf "%g"/"%d" are C's, so the files are byte-identical to
// what the reference writes; Scanf's format " t: %g " (a blank matches any run of white space,
// including none) is mirrored by a small scanner.  This is the interchange format either side
// of the pair loops; nothing here touches the GPU.
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.hpp"

namespace {
using namespace nbh;

struct Scanner {
  std::vector<char> buf;
  size_t pos = 0;
  std::string err;

  int open(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(NBK_ERR_ARG, std::string("cannot open ") + path + ": " + std::strerror(errno));
    char tmp[1 << 16];
    size_t k;
    while ((k = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + k);
    std::fclose(f);
    buf.push_back('\0');
    return 0;
  }
  void blanks() {
    while (buf[pos] && std::isspace((unsigned char)buf[pos])) ++pos;
  }
  bool at_end() {
    blanks();
    return buf[pos] == '\0';
  }
  // The literal part of a Scanf format: blanks in `lit` match any white space, other
  // characters must match exactly.
  bool lit(const char *l) {
    for (; *l; ++l) {
      if (*l == ' ') {
        blanks();
      } else {
        if (buf[pos] != *l) return false;
        ++pos;
      }
    }
    return true;
  }
  bool integer(long *out) {
    blanks();
    char *end = nullptr;
    errno = 0;
    long v = std::strtol(&buf[pos], &end, 10);
    if (end == &buf[pos] || errno) return false;
    pos = (size_t)(end - &buf[0]);
    *out = v;
    return true;
  }
  bool real(double *out) {
    blanks();
    char *end = nullptr;
    double v = std::strtod(&buf[pos], &end);
    if (end == &buf[pos]) return false;
    pos = (size_t)(end - &buf[0]);
    *out = v;
    return true;
  }
  int bad(const char *what) {
    return fail(NBK_ERR_ARG, std::string("snapshot: expected ") + what + " at byte " +
                                 std::to_string(pos) + " (Scanf.Scan_failure in the reference)");
  }
};

// " --- !Particle id: %d " " t: %g " " r: - %g - %g - %g " " v: - %g - %g - %g " " m: %g "
int read_common(Scanner &s, long *id, double *t, double *q, double *v, double *m) {
  if (!s.lit(" --- !Particle id: ") || !s.integer(id)) return s.bad("'--- !Particle id: <int>'");
  if (!s.lit(" t: ") || !s.real(t)) return s.bad("'t: <float>'");
  if (!s.lit(" r: - ") || !s.real(&q[0]) || !s.lit(" - ") || !s.real(&q[1]) || !s.lit(" - ") ||
      !s.real(&q[2]))
    return s.bad("'r: - x - y - z'");
  if (!s.lit(" v: - ") || !s.real(&v[0]) || !s.lit(" - ") || !s.real(&v[1]) || !s.lit(" - ") ||
      !s.real(&v[2]))
    return s.bad("'v: - vx - vy - vz'");
  if (!s.lit(" m: ") || !s.real(m)) return s.bad("'m: <float>'");
  return 0;
}

void print_common(FILE *f, int id, double t, const double *q, const double *v, double m) {
  std::fprintf(f, "--- !Particle\n");
  std::fprintf(f, "id: %d\n", id);
  std::fprintf(f, "t: %g\n", t);
  std::fprintf(f, "r:\n");
  std::fprintf(f, "  - %g\n  - %g\n  - %g\n", q[0], q[1], q[2]);
  std::fprintf(f, "v:\n");
  std::fprintf(f, "  - %g\n  - %g\n  - %g\n", v[0], v[1], v[2]);
  std::fprintf(f, "m: %g\n", m);
}

FILE *open_out(const char *path, int append) {
  FILE *f = std::fopen(path, append ? "ab" : "wb");
  if (!f) fail(NBK_ERR_ARG, std::string("cannot open ") + path + ": " + std::strerror(errno));
  return f;
}
}  // namespace

extern "C" {

int nbh_advancer_print(const char *path, int append, int n, const int *id, const double *t,
                       const double *m, const double *q, const double *p) {
  if (!path || n < 0 || (n && (!id || !t || !m || !q || !p))) return fail(NBK_ERR_ARG, "nbh_advancer_print: bad arguments");
  FILE *f = open_out(path, append);
  if (!f) return NBK_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    const double v[3] = {p[3 * i] / m[i], p[3 * i + 1] / m[i], p[3 * i + 2] / m[i]};  // advancer.ml:95
    print_common(f, id[i], t[i], q + 3 * (size_t)i, v, m[i]);
  }
  std::fclose(f);
  return 0;
}

int nbh_advancer_read(const char *path, int cap, int *n, int *id, double *t, double *m, double *q,
                      double *p) {
  if (!path || !n || cap < 0) return fail(NBK_ERR_ARG, "nbh_advancer_read: bad arguments");
  Scanner s;
  NBH_TRY(s.open(path));
  int k = 0;
  while (!s.at_end()) {
    long idv;
    double tv, qv[3], vv[3], mv;
    NBH_TRY(read_common(s, &idv, &tv, qv, vv, &mv));
    if (k < cap) {
      id[k] = (int)idv;
      t[k] = tv;
      m[k] = mv;
      for (int c = 0; c < 3; ++c) {
        q[3 * k + c] = qv[c];
        p[3 * k + c] = vv[c] * mv;  // Array.map (fun v -> v *. m) v, advancer.ml:124
      }
    }
    ++k;
  }
  *n = k;
  if (k > cap) return fail(NBK_ERR_ARG, "nbh_advancer_read: file holds " + std::to_string(k) + " bodies, capacity " + std::to_string(cap));
  return 0;
}

int nbh_leapfrog_print(const char *path, int append, int n, const int *id, const double *t,
                       const double *dt_max, const double *m, const double *q, const double *v) {
  if (!path || n < 0 || (n && (!id || !t || !dt_max || !m || !q || !v))) return fail(NBK_ERR_ARG, "nbh_leapfrog_print: bad arguments");
  FILE *f = open_out(path, append);
  if (!f) return NBK_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    print_common(f, id[i], t[i], q + 3 * (size_t)i, v + 3 * (size_t)i, m[i]);
    std::fprintf(f, "t_max: %g\n", t[i] + dt_max[i]);  // leapfrog.ml:53
  }
  std::fclose(f);
  return 0;
}

int nbh_leapfrog_read(const char *path, int cap, int *n, int *id, double *t, double *dt_max,
                      double *m, double *q, double *v) {
  if (!path || !n || cap < 0) return fail(NBK_ERR_ARG, "nbh_leapfrog_read: bad arguments");
  Scanner s;
  NBH_TRY(s.open(path));
  int k = 0;
  while (!s.at_end()) {
    long idv;
    double tv, qv[3], vv[3], mv, tm;
    NBH_TRY(read_common(s, &idv, &tv, qv, vv, &mv));
    if (!s.lit(" t_max: ") || !s.real(&tm)) return s.bad("'t_max: <float>'");
    if (k < cap) {
      id[k] = (int)idv;
      t[k] = tv;
      m[k] = mv;
      dt_max[k] = tm - tv;  // leapfrog.ml:62
      for (int c = 0; c < 3; ++c) {
        q[3 * k + c] = qv[c];
        v[3 * k + c] = vv[c];
      }
    }
    ++k;
  }
  *n = k;
  if (k > cap) return fail(NBK_ERR_ARG, "nbh_leapfrog_read: file holds " + std::to_string(k) + " bodies, capacity " + std::to_string(cap));
  return 0;
}

}  // extern "C"
