// Drop-in demonstration: the pipeline of /root/reference/examples/embedder.cpp:213-228 (Galerkin
// coarse graphs -> timed partition::embed -> NaN check) written against the reference's own
// interface, compiled against graph-embed_b200's headers and linked with libgraphembed_b200.so.
// The only source-level difference from a reference build is the include directory.
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>

#include "embed.hpp"
#include "export.hpp"

namespace {

// nx x ny 4-neighbour grid (BASELINE config 1 uses 100 x 100)
SparseMatrix grid(int nx, int ny) {
  linalgcpp::CooMatrix<double> coo(nx * ny, nx * ny);
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      const int v = x * ny + y;
      if (x + 1 < nx) { coo.Add(v, v + ny, 1.0); coo.Add(v + ny, v, 1.0); }
      if (y + 1 < ny) { coo.Add(v, v + 1, 1.0); coo.Add(v + 1, v, 1.0); }
    }
  return coo.ToSparse();
}

// Aggregates of 2 x 2 cells: a stand-in for partition::partition, which stays on the host and
// is an input to the hot path.
SparseMatrix blocks(int nx, int ny) {
  const int bx = (nx + 1) / 2, by = (ny + 1) / 2;
  linalgcpp::CooMatrix<double> coo(bx * by, nx * ny);
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) coo.Add((x / 2) * by + (y / 2), x * ny + y, 1.0);
  return coo.ToSparse();
}

}  // namespace

// `--gpus N`: partition::forceAtlas (include/forceatlas.hpp:89-305) on a 200 x 200 grid, sharded
// over N GPUs of this box by the library (one process, NCCL inside the context).
int flat_on_gpus(int gpus, int dimension) {
  ge_b200::options().gpus = gpus;
  ge_b200::options().seed = 7;
  const SparseMatrix A = grid(200, 200);
  std::vector<std::vector<double>> coords(0);
  linalgcpp::Timer timer(linalgcpp::Timer::Start::True);
  partition::forceAtlas(A, dimension, coords, 20);
  timer.Click();
  std::cout << "forceAtlas: " << A.Rows() << " vertices, 20 iterations on "
            << ge_context_device_count(ge_b200::default_context()) << " GPU(s) in " << timer[0] << "s"
            << std::endl;
  double sum = 0.0;
  for (size_t i = 0; i < coords.size(); i++)
    for (int k = 0; k < dimension; k++) {
      if (std::isnan(coords[i][k])) return 1;
      sum += coords[i][k];
    }
  std::cout.precision(17);
  std::cout << "checksum " << sum << std::endl;
  return 0;
}

int main(int argc, char** argv) {
  std::string outdir;  // --out DIR: write the layout and the plot inputs like examples/embedder.cpp:230-289
  for (int a = 1; a + 1 < argc; ++a) {
    if (std::string(argv[a]) == "--gpus") return flat_on_gpus(std::atoi(argv[a + 1]), 2);
    if (std::string(argv[a]) == "--out") outdir = argv[a + 1];
  }
  int nx = argc > 1 ? std::atoi(argv[1]) : 64, ny = nx;
  const int dimension = argc > 2 ? std::atoi(argv[2]) : 2;
  std::vector<SparseMatrix> As = {grid(nx, ny)}, hierarchy;
  while (nx * ny > 40) {
    hierarchy.push_back(blocks(nx, ny));
    const SparseMatrix& P = hierarchy.back();
    As.push_back(P.Mult(As.back()).Mult(P.Transpose()));  // examples/embedder.cpp:215
    // the same product on the device (ge_galerkin) must give the same matrix, entry for entry
    const SparseMatrix G = ge_b200::galerkin(As[As.size() - 2], P);
    if (G.GetIndptr() != As.back().GetIndptr() || G.GetIndices() != As.back().GetIndices() ||
        G.GetData() != As.back().GetData()) {
      std::cerr << "device Galerkin product differs from P.Mult(A).Mult(P.Transpose())" << std::endl;
      return 3;
    }
    nx = (nx + 1) / 2;
    ny = (ny + 1) / 2;
  }
  std::cout << "levels:";
  for (const auto& A : As) std::cout << " " << A.Rows();
  std::cout << std::endl << "starting embedding: " << std::endl;
  linalgcpp::Timer timer(linalgcpp::Timer::Start::True);
  std::vector<std::vector<double>> coords = partition::embed(As, hierarchy, dimension);
  timer.Click();
  std::cout << "embedded! in time " << timer[0] << "s" << std::endl;
  for (size_t i = 0; i < coords.size(); i++)
    for (int k = 0; k < dimension; k++)
      if (std::isnan(coords[i][k])) {  // examples/embedder.cpp:224-228
        std::cerr << "NaN at vertex " << i << std::endl;
        return 1;
      }
  // the plugin interface (include/embed.hpp:40-49) with the B200 solver as the functor
  std::vector<std::vector<double>> via =
      partition::embedVia(As, hierarchy, dimension, partition::forceAtlasMultilevelEmbedder());
  if (via.size() != coords.size()) return 2;
  std::cout << "embedVia ok: " << via.size() << " vertices" << std::endl;
  if (!outdir.empty()) {
    partition::writeCoords(coords, outdir + "/coords.txt");  // include/export.hpp:23
    ge_b200::writePlotInputs(As[0], hierarchy, coords, dimension, outdir + "/part.temp",
                             outdir + "/coords.temp", outdir + "/mat.temp");
    std::cout << "wrote " << outdir << "/{coords.txt,part.temp,coords.temp,mat.temp}" << std::endl;
  }
  return 0;
}
