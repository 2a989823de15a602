// .npy files either side of the hot path: the per-frame cube score files
// `cube_feat/%06d.npy` ([6,1000,7,7] float32, written at static_model/dataset_feat_extractor.py:187-189,
// read at temporal_model/test_temporal.py:64,70 and data/dataset.py:65) and the equirectangular
// result maps `%05d.npy` ([14,28] float32, temporal_model/test_temporal.py:86-88).
//
// Reader: format versions 1.0 / 2.0 / 3.0, little-endian C-order arrays of f4 / f8 / f2 / u1 / i4 / i8,
// converted to float32 on the way into a caller-owned (ideally pinned) host buffer — the same
// conversion the reference's torch.FloatTensor(np.load(...)) performs (test_temporal.py:70-78).
// Writer: byte-identical to numpy.save(arr) of a C-contiguous float32 array (version 1.0 header,
// padded to a multiple of 64 bytes), so files are interchangeable with the reference's.
#include <ctype.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/cp360.h"

namespace cp360 {
void set_error(const char* fmt, ...);
}
using cp360::set_error;

namespace {

struct NpyHeader {
  std::string descr;
  bool fortran = false;
  std::vector<int64_t> shape;
  int64_t data_offset = 0;
};

int parse_header(FILE* fp, const char* path, NpyHeader* h) {
  unsigned char pre[12];
  if (fread(pre, 1, 10, fp) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) {
    set_error("%s: not a .npy file", path);
    return CP360_ERR_BAD_ARG;
  }
  const int major = pre[6];
  if (pre[7] != 0) { set_error("%s: unsupported .npy version %d.%d", path, major, (int)pre[7]); return CP360_ERR_BAD_ARG; }
  size_t hlen = 0, pre_len = 10;
  if (major == 1) {
    hlen = (size_t)pre[8] | ((size_t)pre[9] << 8);
  } else if (major == 2 || major == 3) {
    if (fread(pre + 10, 1, 2, fp) != 2) { set_error("%s: truncated header", path); return CP360_ERR_BAD_ARG; }
    hlen = (size_t)pre[8] | ((size_t)pre[9] << 8) | ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24);
    pre_len = 12;
  } else {
    set_error("%s: unsupported .npy version %d", path, major);
    return CP360_ERR_BAD_ARG;
  }
  if (hlen == 0 || hlen > (1u << 20)) { set_error("%s: implausible header length", path); return CP360_ERR_BAD_ARG; }
  std::string hdr(hlen, '\0');
  if (fread(&hdr[0], 1, hlen, fp) != hlen) { set_error("%s: truncated header", path); return CP360_ERR_BAD_ARG; }
  h->data_offset = (int64_t)(pre_len + hlen);

  // The header is a Python dict literal; numpy evaluates it with ast.literal_eval and insists on exactly the keys
  // descr / fortran_order / shape. Parse that grammar strictly (anything numpy would reject is rejected here):
  //   '{' entry (',' entry)* [','] '}' spaces ['\n']      entry := str ':' (str | True | False | '(' ints ')')
  size_t i = 0;
  const size_t n = hdr.size();
  auto ws = [&] { while (i < n && (hdr[i] == ' ' || hdr[i] == '\t' || hdr[i] == '\n')) ++i; };
  auto bad = [&](const char* what) { set_error("%s: malformed header (%s)", path, what); return CP360_ERR_BAD_ARG; };
  auto quoted = [&](std::string* out) -> bool {
    if (i >= n || (hdr[i] != '\'' && hdr[i] != '"')) return false;
    const char q = hdr[i++];
    const size_t b = i;
    while (i < n && hdr[i] != q) {
      const unsigned char c = (unsigned char)hdr[i];
      if (c < 0x20 || c > 0x7e || c == '\\') return false;     // plain printable ASCII only (no escapes in dtype strings)
      ++i;
    }
    if (i >= n) return false;
    *out = hdr.substr(b, i - b);
    ++i;
    return true;
  };
  bool have_descr = false, have_fortran = false, have_shape = false;
  ws();
  if (i >= n || hdr[i] != '{') return bad("no dict");
  ++i;
  for (;;) {
    ws();
    if (i < n && hdr[i] == '}') { ++i; break; }
    std::string key;
    if (!quoted(&key)) return bad("key");
    ws();
    if (i >= n || hdr[i] != ':') return bad("':'");
    ++i;
    ws();
    if (key == "descr") {
      if (have_descr || !quoted(&h->descr)) return bad("'descr' (structured dtypes are not supported)");
      have_descr = true;
    } else if (key == "fortran_order") {
      if (have_fortran) return bad("duplicate key");
      if (hdr.compare(i, 4, "True") == 0) { h->fortran = true; i += 4; }
      else if (hdr.compare(i, 5, "False") == 0) { h->fortran = false; i += 5; }
      else return bad("'fortran_order'");
      if (i < n && (isalnum((unsigned char)hdr[i]) || hdr[i] == '_')) return bad("'fortran_order'");
      have_fortran = true;
    } else if (key == "shape") {
      if (have_shape || i >= n || hdr[i] != '(') return bad("'shape'");
      ++i;
      int count = 0;
      bool comma = true;                                        // a number may follow '(' or ','
      for (;;) {
        ws();
        if (i < n && hdr[i] == ')') { ++i; break; }
        if (!comma || i >= n || !isdigit((unsigned char)hdr[i])) return bad("'shape'");
        int64_t v = 0;
        const size_t b = i;
        while (i < n && isdigit((unsigned char)hdr[i])) {
          if (v > (INT64_MAX - 9) / 10) return bad("'shape' extent too large");
          v = v * 10 + (hdr[i++] - '0');
        }
        if (i - b > 1 && hdr[b] == '0') return bad("'shape'");  // leading zeros are a Python syntax error
        if (i < n && hdr[i] == 'L') ++i;                        // python-2 era long suffix
        h->shape.push_back(v);
        ++count;
        ws();
        comma = i < n && hdr[i] == ',';
        if (comma) ++i;
      }
      if (count == 1 && !comma) return bad("'shape' is not a tuple");   // "(5)" is an int, numpy rejects it
      have_shape = true;
    } else {
      return bad("unexpected key");
    }
    ws();
    if (i < n && hdr[i] == ',') { ++i; continue; }
    ws();
    if (i < n && hdr[i] == '}') { ++i; break; }
    return bad("',' or '}'");
  }
  while (i < n && hdr[i] == ' ') ++i;                        // numpy pads with spaces and ends the header with one '\n'
  if (i < n && hdr[i] == '\n') ++i;
  if (i != n) return bad("trailing characters");
  if (!have_descr || !have_fortran || !have_shape) return bad("missing key");
  return CP360_OK;
}

// IEEE binary16 -> binary32 (exact)
float half_to_float(uint16_t hbits) {
  const uint32_t sign = (uint32_t)(hbits & 0x8000u) << 16;
  uint32_t exp = (hbits >> 10) & 0x1fu, man = hbits & 0x3ffu, out;
  if (exp == 0) {
    if (man == 0) {
      out = sign;
    } else {
      exp = 127 - 15 + 1;
      while (!(man & 0x400u)) { man <<= 1; --exp; }
      out = sign | (exp << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    out = sign | 0x7f800000u | (man << 13);
  } else {
    out = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &out, 4);
  return f;
}

int elem_size_of(const std::string& d) {
  if (d == "<f4" || d == "=f4" || d == "<i4" || d == "=i4") return 4;
  if (d == "<f8" || d == "=f8" || d == "<i8" || d == "=i8") return 8;
  if (d == "<f2" || d == "=f2") return 2;
  if (d == "|u1" || d == "<u1" || d == "=u1") return 1;
  return 0;
}

}  // namespace

extern "C" {

int cp360_npy_read_header(const char* path, char* descr, int descr_len, int* ndim, int64_t* shape,
                          int max_dims, int64_t* data_offset, int* fortran_order) {
  if (!path) { set_error("npy: null path"); return CP360_ERR_BAD_ARG; }
  FILE* fp = fopen(path, "rb");
  if (!fp) { set_error("%s: %s", path, strerror(errno)); return CP360_ERR_BAD_ARG; }
  NpyHeader h;
  const int rc = parse_header(fp, path, &h);
  fclose(fp);
  if (rc != CP360_OK) return rc;
  if (descr && descr_len > 0) snprintf(descr, (size_t)descr_len, "%s", h.descr.c_str());
  if (ndim) *ndim = (int)h.shape.size();
  if (shape) {
    if ((int)h.shape.size() > max_dims) { set_error("%s: %d dims > %d", path, (int)h.shape.size(), max_dims); return CP360_ERR_RANGE; }
    for (size_t i = 0; i < h.shape.size(); ++i) shape[i] = h.shape[i];
  }
  if (data_offset) *data_offset = h.data_offset;
  if (fortran_order) *fortran_order = h.fortran ? 1 : 0;
  return CP360_OK;
}

int cp360_npy_read_f32(const char* path, float* dst_host, int64_t n_elems) {
  if (!path || (!dst_host && n_elems > 0) || n_elems < 0) { set_error("npy: bad argument"); return CP360_ERR_BAD_ARG; }
  FILE* fp = fopen(path, "rb");
  if (!fp) { set_error("%s: %s", path, strerror(errno)); return CP360_ERR_BAD_ARG; }
  NpyHeader h;
  int rc = parse_header(fp, path, &h);
  if (rc != CP360_OK) { fclose(fp); return rc; }
  int64_t n = 1;
  for (int64_t v : h.shape) {
    if (v != 0 && n > INT64_MAX / v) { fclose(fp); set_error("%s: element count overflows", path); return CP360_ERR_RANGE; }
    n *= v;
  }
  const int es = elem_size_of(h.descr);
  if (h.fortran && h.shape.size() > 1) { fclose(fp); set_error("%s: fortran_order arrays are not supported", path); return CP360_ERR_SHAPE; }
  if (es == 0) { fclose(fp); set_error("%s: unsupported dtype '%s'", path, h.descr.c_str()); return CP360_ERR_BAD_ARG; }
  if (n != n_elems) { fclose(fp); set_error("%s: holds %lld elements, caller expects %lld", path, (long long)n, (long long)n_elems); return CP360_ERR_SHAPE; }
  const char kind = h.descr[1];
  bool ok = true;
  if (kind == 'f' && es == 4) {
    ok = fread(dst_host, 4, (size_t)n, fp) == (size_t)n;
  } else {
    const size_t chunk = 1 << 16;
    std::vector<unsigned char> buf(chunk * (size_t)es);
    for (int64_t done = 0; done < n && ok;) {
      const size_t m = (size_t)((n - done) < (int64_t)chunk ? (n - done) : (int64_t)chunk);
      ok = fread(buf.data(), (size_t)es, m, fp) == m;
      if (!ok) break;
      float* d = dst_host + done;
      if (kind == 'f' && es == 8) { const double* s = (const double*)buf.data(); for (size_t i = 0; i < m; ++i) d[i] = (float)s[i]; }
      else if (kind == 'f' && es == 2) { const uint16_t* s = (const uint16_t*)buf.data(); for (size_t i = 0; i < m; ++i) d[i] = half_to_float(s[i]); }
      else if (kind == 'u') { const uint8_t* s = buf.data(); for (size_t i = 0; i < m; ++i) d[i] = (float)s[i]; }
      else if (kind == 'i' && es == 4) { const int32_t* s = (const int32_t*)buf.data(); for (size_t i = 0; i < m; ++i) d[i] = (float)s[i]; }
      else { const int64_t* s = (const int64_t*)buf.data(); for (size_t i = 0; i < m; ++i) d[i] = (float)s[i]; }
      done += (int64_t)m;
    }
  }
  fclose(fp);
  if (!ok) { set_error("%s: truncated data", path); return CP360_ERR_BAD_ARG; }
  return CP360_OK;
}

int cp360_npy_write_f32(const char* path, const float* src_host, int ndim, const int64_t* shape) {
  if (!path || ndim < 0 || ndim > 32 || (ndim > 0 && !shape)) { set_error("npy: bad argument"); return CP360_ERR_BAD_ARG; }
  int64_t n = 1;
  std::string shp = "(";
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0) { set_error("npy: negative extent"); return CP360_ERR_BAD_ARG; }
    n *= shape[i];
    shp += std::to_string((long long)shape[i]);
    if (ndim == 1) shp += ",";
    else if (i + 1 < ndim) shp += ", ";
  }
  shp += ")";
  if (n > 0 && !src_host) { set_error("npy: null data"); return CP360_ERR_BAD_ARG; }
  // numpy.lib.format: dict literal, then spaces so that magic+len+header is a multiple of 64, '\n' last
  std::string hdr = "{'descr': '<f4', 'fortran_order': False, 'shape': " + shp + ", }";
  // numpy's _write_array_header: spare room so the first extent can grow in place (21 digits), then
  // _wrap_header: padlen = 64 - ((magic + len field + header + '\n') % 64), i.e. 1..64 spaces, never 0
  if (ndim > 0) {
    const size_t digits = std::to_string((long long)shape[0]).size();
    if (digits < 21) hdr.append(21 - digits, ' ');
  }
  const size_t hlen = hdr.size() + 1;
  hdr.append(64 - ((10 + hlen) % 64), ' ');
  hdr.push_back('\n');
  if (hdr.size() > 65535) { set_error("npy: header too long for format 1.0"); return CP360_ERR_RANGE; }
  const std::string tmp = std::string(path) + ".tmp~";
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) { set_error("%s: %s", tmp.c_str(), strerror(errno)); return CP360_ERR_BAD_ARG; }
  const unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hdr.size() & 0xff),
                                 (unsigned char)(hdr.size() >> 8)};
  bool ok = fwrite(pre, 1, 10, fp) == 10 && fwrite(hdr.data(), 1, hdr.size(), fp) == hdr.size();
  if (ok && n > 0) ok = fwrite(src_host, 4, (size_t)n, fp) == (size_t)n;
  ok = (fclose(fp) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    set_error("%s: write failed: %s", path, strerror(errno));
    remove(tmp.c_str());
    return CP360_ERR_BAD_ARG;
  }
  return CP360_OK;
}

}  // extern "C"
