// Data ingest for the scan path: the delimited-text matrices the reference reads
// with readdlm — readBXDpheno / readBXDgeno (src/readData.jl:159-165), readGenoProb_ExcludeComplements
// (src/readData.jl:85-96) — parsed by host threads straight into the column-major Float64 layout the scans take,
// in host memory or (one H2D copy from pinned staging) in device memory, so that G and Y need not pass through
// Julia arrays at all.  Host-only code: no kernels here.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "../../include/blmm_b200.h"

namespace {

struct Parsed {
  int64_t rows = 0, cols = 0;
  std::vector<double> data;  // column-major rows x cols
  std::string err;
};

// one field -> double; empty / non-numeric fields (NA, strings) become NaN like a failed Float64 conversion would
// be an error in the reference: we report them instead of guessing
bool parse_field(const char* b, const char* e, double* out) {
  while (b < e && (*b == ' ' || *b == '\t' || *b == '"')) ++b;
  while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r' || e[-1] == '"')) --e;
  if (b == e) return false;
  if (*b == '+') ++b;
  auto r = std::from_chars(b, e, *out);
  return r.ec == std::errc() && r.ptr == e;
}

Parsed parse_csv(const char* path, char delim, int64_t skip_rows, int64_t first_col, int64_t col_step,
                 int64_t drop_last_cols) {
  Parsed P;
  FILE* f = fopen(path, "rb");
  if (!f) {
    P.err = std::string("cannot open ") + path;
    return P;
  }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::string buf((size_t)sz, '\0');
  if (sz > 0 && fread(&buf[0], 1, (size_t)sz, f) != (size_t)sz) {
    fclose(f);
    P.err = std::string("short read on ") + path;
    return P;
  }
  fclose(f);
  // line starts (blank trailing lines ignored)
  std::vector<size_t> ls;
  size_t pos = 0;
  while (pos < buf.size()) {
    size_t nl = buf.find('\n', pos);
    if (nl == std::string::npos) nl = buf.size();
    bool blank = true;
    for (size_t i = pos; i < nl; ++i)
      if (buf[i] != ' ' && buf[i] != '\r' && buf[i] != '\t') {
        blank = false;
        break;
      }
    if (!blank) ls.push_back(pos);
    pos = nl + 1;
  }
  if ((int64_t)ls.size() <= skip_rows) {
    P.err = "no data rows after the skipped header";
    return P;
  }
  const int64_t rows = (int64_t)ls.size() - skip_rows;
  auto line_end = [&](size_t start) {
    size_t nl = buf.find('\n', start);
    return nl == std::string::npos ? buf.size() : nl;
  };
  // fields of the first data row fix the column count
  int64_t nfield = 1;
  {
    const size_t b = ls[(size_t)skip_rows], e = line_end(b);
    for (size_t i = b; i < e; ++i) nfield += (buf[i] == delim);
  }
  const int64_t last = nfield - drop_last_cols;  // exclusive
  if (first_col < 0 || col_step < 1 || drop_last_cols < 0 || first_col >= last) {
    P.err = "column selection is empty";
    return P;
  }
  const int64_t cols = (last - first_col + col_step - 1) / col_step;
  P.rows = rows;
  P.cols = cols;
  P.data.assign((size_t)rows * cols, 0.0);
  std::atomic<int64_t> bad_row{-1}, bad_col{-1};
  const unsigned hc = std::thread::hardware_concurrency();
  const int W = (int)std::max(1u, std::min(16u, hc ? hc : 1u));
  std::vector<std::thread> th;
  for (int w = 0; w < W; ++w)
    th.emplace_back([&, w]() {
      for (int64_t r = rows * w / W; r < rows * (w + 1) / W; ++r) {
        const size_t b = ls[(size_t)(skip_rows + r)], e = line_end(b);
        int64_t field = 0, c = 0;
        size_t fb = b;
        for (size_t i = b; i <= e; ++i) {
          if (i == e || buf[i] == delim) {
            if (field >= first_col && field < last && (field - first_col) % col_step == 0) {
              double v;
              if (!parse_field(buf.data() + fb, buf.data() + i, &v)) {
                bad_row.store(r);
                bad_col.store(field);
                return;
              }
              P.data[(size_t)c * rows + r] = v;
              ++c;
            }
            ++field;
            fb = i + 1;
          }
        }
        if (c != cols || field != nfield) {  // fewer OR more fields than the first data row: readdlm would not build a matrix
          bad_row.store(r);
          bad_col.store(-2);
          return;
        }
      }
    });
  for (auto& t : th) t.join();
  if (bad_row.load() >= 0) {
    const int64_t r = bad_row.load(), c = bad_col.load();
    P.err = (c == -2) ? "row " + std::to_string(skip_rows + r + 1) + " has a different number of fields"
                      : "non-numeric field at row " + std::to_string(skip_rows + r + 1) + ", column " + std::to_string(c + 1);
    P.data.clear();
  }
  return P;
}

thread_local std::string g_io_err;

}  // namespace

extern "C" {

BLMM_API const char* blmm_io_last_error(void) { return g_io_err.c_str(); }

BLMM_API int blmm_read_csv(const char* path, char delim, int64_t skip_rows, int64_t first_col, int64_t col_step,
                           int64_t drop_last_cols, int device, int64_t* rows_out, int64_t* cols_out, double** data_out) {
  g_io_err.clear();
  if (!path || !rows_out || !cols_out || !data_out) {
    g_io_err = "blmm_read_csv: NULL argument";
    return BLMM_E_INVALID;
  }
  Parsed P;
  try {
    P = parse_csv(path, delim ? delim : ',', skip_rows, first_col, col_step, drop_last_cols);
  } catch (const std::exception& ex) {
    g_io_err = ex.what();
    return BLMM_E_INVALID;
  }
  if (!P.err.empty()) {
    g_io_err = P.err;
    return BLMM_E_INVALID;
  }
  const size_t bytes = P.data.size() * sizeof(double);
  double* out = nullptr;
  if (device < 0) {
    out = (double*)malloc(bytes ? bytes : 8);
    if (!out) {
      g_io_err = "out of host memory";
      return BLMM_E_INVALID;
    }
    memcpy(out, P.data.data(), bytes);
  } else {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&out, bytes ? bytes : 8);
    if (e == cudaSuccess) e = cudaMemcpy(out, P.data.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (out) cudaFree(out);
      g_io_err = std::string("blmm_read_csv: ") + cudaGetErrorString(e);
      return BLMM_E_CUDA;
    }
  }
  *rows_out = P.rows;
  *cols_out = P.cols;
  *data_out = out;
  return BLMM_OK;
}

BLMM_API void blmm_free_matrix(double* data, int device) {
  if (!data) return;
  if (device < 0) {
    free(data);
  } else {
    cudaSetDevice(device);
    cudaFree(data);
  }
}

}  // extern "C"
