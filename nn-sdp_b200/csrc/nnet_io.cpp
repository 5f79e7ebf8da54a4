// .nnet reader (host only): the file format of the reference's fixtures (bench/rand/*.nnet), as read by
// exts/nnet_parser.jl:35-131 (NNet(file)) and turned into a FeedFwdNet by loadFromNnet
// (src/MyNeuralNetwork/network_files.jl): header lines starting with "//", then
//   numLayers, inputSize, outputSize, maxLayerSize
//   layerSizes[0..numLayers]
//   one unused line, mins, maxes, means, ranges
//   per layer: one line per output neuron with its weights, then one line per output neuron with its bias.
// Ms[k] is written column-major as [W_k b_k] (xdims[k+1] x (xdims[k]+1)), the layout nnsdp_net_upload takes.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "internal.h"

using namespace nnsdp;

namespace {

bool read_line(FILE* f, std::string* out) {
  out->clear();
  int c;
  while ((c = fgetc(f)) != EOF) {
    if (c == '\n') return true;
    if (c != '\r') out->push_back((char)c);
  }
  return !out->empty();
}

// comma separated numbers; empty fields are skipped
bool parse_numbers(const std::string& line, std::vector<double>* v) {
  v->clear();
  const char* p = line.c_str();
  while (*p) {
    while (*p == ',' || *p == ' ' || *p == '\t') ++p;
    if (!*p) break;
    char* end = nullptr;
    const double x = strtod(p, &end);
    if (end == p) return false;
    v->push_back(x);
    p = end;
  }
  return true;
}

}  // namespace

extern "C" int32_t nnsdp_nnet_read(const char* path, int64_t max_layers, int64_t* K_out, int64_t* xdims_out,
                                   int64_t max_doubles, double* Ms_out, int64_t* doubles_needed) {
  NN_CHECK(path && K_out, NNSDP_ERR_ARG, "NULL argument");
  FILE* f = fopen(path, "r");
  NN_CHECK(f != nullptr, NNSDP_ERR_ARG, "cannot open %s", path);
  struct Closer {
    FILE* f;
    ~Closer() { fclose(f); }
  } closer{f};
  std::string line;
  std::vector<double> v;
  do {
    NN_CHECK(read_line(f, &line), NNSDP_ERR_ARG, "%s: unexpected end of file in the header", path);
  } while (line.size() >= 2 && line[0] == '/' && line[1] == '/');
  NN_CHECK(parse_numbers(line, &v) && v.size() >= 3, NNSDP_ERR_ARG, "%s: bad architecture line", path);
  const int64_t K = (int64_t)v[0];
  NN_CHECK(K >= 1 && K < 32768, NNSDP_ERR_ARG, "%s: bad number of layers", path);
  NN_CHECK(read_line(f, &line) && parse_numbers(line, &v) && (int64_t)v.size() >= K + 1, NNSDP_ERR_ARG,
           "%s: bad layer-size line", path);
  std::vector<int64_t> xd(K + 1);
  for (int64_t i = 0; i <= K; ++i) {
    xd[i] = (int64_t)v[i];
    NN_CHECK(xd[i] >= 1, NNSDP_ERR_ARG, "%s: bad layer size", path);
  }
  for (int i = 0; i < 5; ++i)  // unused flag, mins, maxes, means, ranges
    NN_CHECK(read_line(f, &line), NNSDP_ERR_ARG, "%s: unexpected end of file in the normalisation block", path);
  int64_t need = 0;
  for (int64_t k = 0; k < K; ++k) need += xd[k + 1] * (xd[k] + 1);
  *K_out = K;
  if (doubles_needed) *doubles_needed = need;
  if (xdims_out) {
    NN_CHECK(max_layers >= K, NNSDP_ERR_ARG, "xdims_out too small: the net has %lld layers", (long long)K);
    for (int64_t i = 0; i <= K; ++i) xdims_out[i] = xd[i];
  }
  if (!Ms_out) return NNSDP_OK;
  NN_CHECK(max_doubles >= need, NNSDP_ERR_ARG, "Ms_out too small: %lld doubles needed", (long long)need);
  double* M = Ms_out;
  for (int64_t k = 0; k < K; ++k) {
    const int64_t nout = xd[k + 1], nin = xd[k];
    for (int64_t i = 0; i < nout; ++i) {
      NN_CHECK(read_line(f, &line) && parse_numbers(line, &v) && (int64_t)v.size() >= nin, NNSDP_ERR_ARG,
               "%s: bad weight row %lld of layer %lld", path, (long long)i, (long long)k);
      for (int64_t j = 0; j < nin; ++j) M[i + j * nout] = v[j];
    }
    for (int64_t i = 0; i < nout; ++i) {
      NN_CHECK(read_line(f, &line) && parse_numbers(line, &v) && !v.empty(), NNSDP_ERR_ARG,
               "%s: bad bias %lld of layer %lld", path, (long long)i, (long long)k);
      M[i + nin * nout] = v[0];
    }
    M += nout * (nin + 1);
  }
  return NNSDP_OK;
}
