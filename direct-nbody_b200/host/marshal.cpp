// marshal.cpp — the reference's binary state streams: files of concatenated
// `Marshal.to_channel chan (bs : B.b array) []` values, written by bin/nbody.ml:80 (states.dat),
// bin/el_mass_segregation.ml:96-108 (checkpoint and states files) and read back by
// Analysis.load_states / load_state (analysis.ml:556-575).  B.b is Advancer.A.body
// (advancer.ml:27-45, 17 fields) in both drivers; Leapfrog.b (leapfrog.ml:5-12, 6 fields) is
// accepted too since Analysis.Make is a functor over BODY.
//
// The byte format is the OCaml runtime's (runtime/caml/intext.h, extern.c, intern.c), restated
// here from its published definition because no OCaml runtime is available to link:
//   header  magic 0x8495A6BE, then four big-endian 32-bit words: data length in bytes, number of
//           shared-able objects, heap words needed on a 32-bit reader, on a 64-bit reader
//           (magic 0x8495A6BF: the "big" header of OCaml >= 4.04 with 64-bit fields, read only)
//   values  0x40|n  small int 0..63      0x00 int8   0x01 int16   0x02 int32   0x03 int64 (big-endian)
//           0x80|tag|(size<<4)  small block (tag < 16, size < 8)
//           0x08 block with 32-bit header word (size << 10 | tag)    0x13 block with 64-bit header
//           0x20|len small string   0x09 string8   0x0A string32
//           0x0B / 0x0C  boxed double, big / little endian
//           0x0D / 0x0E  double array with 8-bit length, big / little endian;  0x0F / 0x07 with 32-bit length
//           0x04 / 0x05 / 0x06  reference to an earlier object, 8 / 16 / 32-bit distance back
// Every block, boxed double, double array and string takes the next object number as it is
// encountered; Marshal preserves physical sharing (the `nan` constants of A.make_body_id,
// advancer.ml:98-116, are ONE boxed float referenced from every fresh body), so the reader keeps
// the object table.  The writer emits no sharing (legal: sharing is an optimisation of the
// encoding; the value read back is structurally equal) and little-endian doubles, like ocamlopt
// on x86-64 / arm64.  Not checked against a real OCaml runtime (none exists in the build image):
// checked by round trips, by hand-assembled streams and by a stream of the reference's own body
// records written by an independent encoder (tests/test_marshal.py), which also reads a stream
// written by the real Marshal module whenever one has been generated with an OCaml toolchain.
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.hpp"

namespace {
using namespace nbh;

enum Kind { K_INT, K_DOUBLE, K_DARRAY, K_BLOCK, K_STRING };

struct Val {
  Kind kind = K_INT;
  long long i = 0;              // K_INT; K_BLOCK: tag
  double d = 0.0;               // K_DOUBLE
  std::vector<double> da;       // K_DARRAY
  std::vector<int> fields;      // K_BLOCK: indices into Reader::pool
  std::string s;                // K_STRING
};

struct Reader {
  const unsigned char *p, *end;
  std::vector<Val> pool;        // every value read
  std::vector<int> objects;     // object number -> pool index (blocks, doubles, arrays, strings)
  std::string err;

  bool need(size_t k) {
    if ((size_t)(end - p) < k) {
      err = "truncated Marshal stream";
      return false;
    }
    return true;
  }
  uint64_t be(int bytes) {
    uint64_t v = 0;
    for (int k = 0; k < bytes; ++k) v = (v << 8) | *p++;
    return v;
  }
  long long sbe(int bytes) {
    uint64_t v = be(bytes);
    if (bytes < 8 && (v >> (8 * bytes - 1))) v |= ~0ULL << (8 * bytes);
    return (long long)v;
  }
  double dbl(bool little) {
    unsigned char b[8];
    for (int k = 0; k < 8; ++k) b[little ? k : 7 - k] = *p++;
    double x;
    std::memcpy(&x, b, 8);  // host is little-endian (x86-64 / arm64)
    return x;
  }
  int add(Val &&v, bool is_object) {
    pool.push_back(std::move(v));
    const int idx = (int)pool.size() - 1;
    if (is_object) objects.push_back(idx);
    return idx;
  }
  int block(long long tag, uint64_t size) {
    if (size > (uint64_t)(end - p)) {  // every field takes at least one byte
      err = "block size exceeds the stream";
      return -1;
    }
    Val v;
    v.kind = K_BLOCK;
    v.i = tag;
    // size-0 blocks are atoms: not entered in the object table (extern.c)
    const int idx = add(std::move(v), size > 0);
    std::vector<int> f((size_t)size);
    for (uint64_t k = 0; k < size; ++k) {
      f[k] = value();
      if (f[k] < 0) return -1;
    }
    pool[idx].fields = std::move(f);
    return idx;
  }
  int darray(uint64_t len, bool little) {
    if (!need(8 * len)) return -1;
    Val v;
    v.kind = K_DARRAY;
    v.da.resize((size_t)len);
    for (uint64_t k = 0; k < len; ++k) v.da[k] = dbl(little);
    return add(std::move(v), true);
  }
  int str(uint64_t len) {
    if (!need(len)) return -1;
    Val v;
    v.kind = K_STRING;
    v.s.assign((const char *)p, (size_t)len);
    p += len;
    return add(std::move(v), true);
  }
  int shared(uint64_t ofs) {
    if (ofs == 0 || ofs > objects.size()) {
      err = "bad shared-object reference";
      return -1;
    }
    return objects[objects.size() - (size_t)ofs];
  }
  int value() {
    if (!need(1)) return -1;
    const unsigned c = *p++;
    if (c >= 0x80) return block(c & 0xF, (c >> 4) & 0x7);
    if (c >= 0x40) {
      Val v;
      v.i = c & 0x3F;
      return add(std::move(v), false);
    }
    if (c >= 0x20) return str(c & 0x1F);
    switch (c) {
      case 0x00: case 0x01: case 0x02: case 0x03: {
        const int bytes = 1 << c;
        if (!need(bytes)) return -1;
        Val v;
        v.i = sbe(bytes);
        return add(std::move(v), false);
      }
      case 0x04: if (!need(1)) return -1; return shared(be(1));
      case 0x05: if (!need(2)) return -1; return shared(be(2));
      case 0x06: if (!need(4)) return -1; return shared(be(4));
      case 0x08: {
        if (!need(4)) return -1;
        const uint64_t h = be(4);
        return block((long long)(h & 0xFF), h >> 10);
      }
      case 0x13: {
        if (!need(8)) return -1;
        const uint64_t h = be(8);
        return block((long long)(h & 0xFF), h >> 10);
      }
      case 0x09: if (!need(1)) return -1; return str(be(1));
      case 0x0A: if (!need(4)) return -1; return str(be(4));
      case 0x0B: case 0x0C: {
        if (!need(8)) return -1;
        Val v;
        v.kind = K_DOUBLE;
        v.d = dbl(c == 0x0C);
        return add(std::move(v), true);
      }
      case 0x0D: case 0x0E: if (!need(1)) return -1; return darray(be(1), c == 0x0E);
      case 0x0F: case 0x07: if (!need(4)) return -1; return darray(be(4), c == 0x07);
      default:
        err = "unsupported Marshal code (closures, custom blocks and code pointers do not occur in "
              "body arrays)";
        return -1;
    }
  }
};

int read_file(const char *path, std::vector<unsigned char> &buf) {
  FILE *f = std::fopen(path, "rb");
  if (!f) return fail(NBK_ERR_ARG, std::string("cannot open ") + path + ": " + std::strerror(errno));
  unsigned char tmp[1 << 16];
  size_t k;
  while ((k = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + k);
  std::fclose(f);
  return 0;
}

// Splits a file of concatenated Marshal values: offsets of each value's data and its length.
struct Chunk {
  size_t data, len;
};
int split_stream(const std::vector<unsigned char> &buf, std::vector<Chunk> &out) {
  size_t pos = 0;
  auto be32 = [&](size_t o) {
    return ((uint64_t)buf[o] << 24) | ((uint64_t)buf[o + 1] << 16) | ((uint64_t)buf[o + 2] << 8) | buf[o + 3];
  };
  while (pos < buf.size()) {
    if (buf.size() - pos < 20) break;  // load_states stops at the first input_value failure
    const uint64_t magic = be32(pos);
    size_t hdr, len;
    if (magic == 0x8495A6BEu) {
      hdr = 20;
      len = (size_t)be32(pos + 4);
    } else if (magic == 0x8495A6BFu) {
      if (buf.size() - pos < 32) break;
      hdr = 32;
      len = (size_t)((be32(pos + 8) << 32) | be32(pos + 12));
    } else {
      return fail(NBK_ERR_ARG, "not a Marshal stream (bad magic number; compressed streams of OCaml "
                               ">= 5.1 are not supported)");
    }
    if (buf.size() - pos - hdr < len) break;  // a truncated last value: ignored like the reference
    out.push_back({pos + hdr, len});
    pos += hdr + len;
  }
  return 0;
}

bool get_double(const Reader &r, int idx, double *x) {
  const Val &v = r.pool[idx];
  if (v.kind != K_DOUBLE) return false;
  *x = v.d;
  return true;
}
bool get_vec3(const Reader &r, int idx, double *x) {
  const Val &v = r.pool[idx];
  if (v.kind != K_DARRAY || v.da.size() != 3) return false;
  for (int k = 0; k < 3; ++k) x[k] = v.da[k];
  return true;
}

// ---- writer -------------------------------------------------------------------------------------
struct Writer {
  std::vector<unsigned char> out;
  uint64_t objects = 0, words32 = 0, words64 = 0;
  void be(uint64_t v, int bytes) {
    for (int k = bytes - 1; k >= 0; --k) out.push_back((unsigned char)(v >> (8 * k)));
  }
  void integer(long long n) {
    if (n >= 0 && n < 0x40) out.push_back((unsigned char)(0x40 | n));
    else if (n >= -128 && n < 128) { out.push_back(0x00); be((uint64_t)n, 1); }
    else if (n >= -32768 && n < 32768) { out.push_back(0x01); be((uint64_t)n, 2); }
    else if (n >= -(1LL << 30) && n < (1LL << 30)) { out.push_back(0x02); be((uint64_t)n, 4); }
    else { out.push_back(0x03); be((uint64_t)n, 8); }
  }
  void block_header(int tag, uint64_t size) {
    if (size == 0) {  // atom
      out.push_back((unsigned char)(0x80 | tag));
      return;
    }
    if (tag < 16 && size < 8) out.push_back((unsigned char)(0x80 | tag | (size << 4)));
    else if (size < (1ULL << 22)) { out.push_back(0x08); be((size << 10) | (uint64_t)tag, 4); }
    else { out.push_back(0x13); be((size << 10) | (uint64_t)tag, 8); }
    ++objects;
    words32 += 1 + size;
    words64 += 1 + size;
  }
  void boxed(double x) {
    out.push_back(0x0C);
    unsigned char b[8];
    std::memcpy(b, &x, 8);
    out.insert(out.end(), b, b + 8);
    ++objects;
    words32 += 1 + 2;
    words64 += 1 + 1;
  }
  void vec3(const double *x) {
    out.push_back(0x0E);
    out.push_back(3);
    for (int k = 0; k < 3; ++k) {
      unsigned char b[8];
      std::memcpy(b, &x[k], 8);
      out.insert(out.end(), b, b + 8);
    }
    ++objects;
    words32 += 1 + 6;
    words64 += 1 + 3;
  }
  int flush(const char *path, int append) {
    if (out.size() >= (1ULL << 32) || words32 >= (1ULL << 32) || words64 >= (1ULL << 32))
      return fail(NBK_ERR_ARG, "state too large for the small Marshal header");
    std::vector<unsigned char> hdr;
    auto be32 = [&](uint64_t v) {
      for (int k = 3; k >= 0; --k) hdr.push_back((unsigned char)(v >> (8 * k)));
    };
    be32(0x8495A6BEu);
    be32(out.size());
    be32(objects);
    be32(words32);
    be32(words64);
    FILE *f = std::fopen(path, append ? "ab" : "wb");
    if (!f) return fail(NBK_ERR_ARG, std::string("cannot open ") + path + ": " + std::strerror(errno));
    const bool ok = std::fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size() &&
                    std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? 0 : fail(NBK_ERR_ARG, std::string("short write to ") + path);
  }
};

}  // namespace

extern "C" {

// Number of states (body arrays) in a stream file: Array.length (Analysis.load_states file).
int nbh_marshal_count_states(const char *path, int *nstates) {
  std::vector<unsigned char> buf;
  NBH_TRY(read_file(path, buf));
  std::vector<Chunk> ch;
  NBH_TRY(split_stream(buf, ch));
  *nstates = (int)ch.size();
  return 0;
}

// Analysis.load_state index file (analysis.ml:567-575) for Advancer.A.body arrays: *n = bodies in
// that state; id[cap], state[cap][35] in the record order of nbh_advancer_advance; the int field
// output_counter is dropped.  kind: out, 17 = Advancer.A.body, 6 = Leapfrog.b (then the 35-double
// row holds {t, m, q[3], v[3], dt_max} in its first 9 slots, the rest NaN).
int nbh_marshal_read_state(const char *path, int index, int cap, int *n, int *kind, int *id,
                           double *state) {
  std::vector<unsigned char> buf;
  NBH_TRY(read_file(path, buf));
  std::vector<Chunk> ch;
  NBH_TRY(split_stream(buf, ch));
  if (index < 0 || index >= (int)ch.size())
    return fail(NBK_ERR_ARG, "load_state: End_of_file (state index beyond the stream)");
  Reader r;
  r.p = buf.data() + ch[index].data;
  r.end = r.p + ch[index].len;
  const int root = r.value();
  if (root < 0) return fail(NBK_ERR_ARG, "load_state: " + r.err);
  const Val &arr = r.pool[root];
  if (arr.kind != K_BLOCK || arr.i != 0) return fail(NBK_ERR_ARG, "load_state: not a body array");
  *n = (int)arr.fields.size();
  if (*n > cap) return fail(NBK_ERR_ARG, "load_state: more bodies than the caller's capacity");
  *kind = 0;
  for (int b = 0; b < *n; ++b) {
    const Val &rec = r.pool[arr.fields[b]];
    double *row = state + (size_t)b * NBH_ADV_STRIDE;
    for (int k = 0; k < NBH_ADV_STRIDE; ++k) row[k] = std::nan("");
    if (rec.kind != K_BLOCK) return fail(NBK_ERR_ARG, "load_state: body is not a record");
    const std::vector<int> &f = rec.fields;
    bool ok = true;
    if (f.size() == 17) {  // Advancer.A.body, advancer.ml:27-45
      *kind = 17;
      ok = r.pool[f[0]].kind == K_INT;
      id[b] = (int)r.pool[f[0]].i;
      ok = ok && get_double(r, f[1], row + 0) && get_double(r, f[2], row + 1) &&
           get_vec3(r, f[3], row + 2) && get_vec3(r, f[4], row + 5) && get_double(r, f[5], row + 8) &&
           get_double(r, f[6], row + 9) && get_double(r, f[7], row + 10);
      for (int a = 0; a < 8 && ok; ++a) ok = get_vec3(r, f[8 + a], row + 11 + 3 * a);
    } else if (f.size() == 6) {  // Leapfrog.b, leapfrog.ml:5-12
      *kind = 6;
      ok = r.pool[f[0]].kind == K_INT;
      id[b] = (int)r.pool[f[0]].i;
      ok = ok && get_double(r, f[1], row + 0) && get_double(r, f[2], row + 8) &&
           get_double(r, f[3], row + 1) && get_vec3(r, f[4], row + 2) && get_vec3(r, f[5], row + 5);
    } else {
      ok = false;
    }
    if (!ok) return fail(NBK_ERR_ARG, "load_state: record is neither Advancer.A.body nor Leapfrog.b");
  }
  return 0;
}

// Marshal.to_channel chan (bs : Advancer.A.body array) [] (bin/nbody.ml:80,
// bin/el_mass_segregation.ml:99,106): one state appended to (or starting) the stream file.
int nbh_marshal_write_state(const char *path, int append, int n, const int *id, const double *state) {
  if (n < 0 || !id || !state) return fail(NBK_ERR_ARG, "Marshal.to_channel: bad arguments");
  Writer w;
  w.block_header(0, (uint64_t)n);
  for (int b = 0; b < n; ++b) {
    const double *row = state + (size_t)b * NBH_ADV_STRIDE;
    w.block_header(0, 17);
    w.integer(id[b]);
    w.boxed(row[0]);   // t
    w.boxed(row[1]);   // m
    w.vec3(row + 2);   // q
    w.vec3(row + 5);   // p
    w.boxed(row[8]);   // h
    w.boxed(row[9]);   // hmax
    w.boxed(row[10]);  // t0
    for (int a = 0; a < 8; ++a) w.vec3(row + 11 + 3 * a);  // dVdq0-2, p0, q0, q1, q2, cs
    w.integer(0);      // output_counter
  }
  return w.flush(path, append);
}

}  // extern "C"
