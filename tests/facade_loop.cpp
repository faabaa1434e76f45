// tests/facade_loop.cpp -- the reference's frame loop with its PER-PARTICLE calls (SoilMachine.cpp:287-320:
// construct a particle, `while (move && interact);`, flood, then the seep pass, then the wind particles, then the
// frequency maps) compiled against include/soilmachine/soilmachine.hpp instead of the reference headers.  Nothing
// in the loop body is adapted to the facade: this is the source-compatible slow path (one particle = one batch of
// one).  The map it leaves is written to a file and compared with the reference run sequentially on the same seed
// (tests/test_gpu_parity.py::test_facade_per_particle_loop_matches_reference).
//   facade_loop <file.soil> <dim> <nwater> <nwind> <frames> <out.bin>
#include <cstdio>
#include <cstdlib>
#include "../include/soilmachine/soilmachine.hpp"
using namespace soilmachine;

int SIZEX = 64, SIZEY = 64, SCALE = 80, SLICE = 160, NWATER = 40, NWIND = 20, SEED = 42;
struct CountingPool {                       // stands in for Vertexpool<Vertex>: receives the mesh
  size_t uploads = 0, nvertices = 0;
  void upload(const float*, size_t n) { uploads++; nvertices = n; }
} vertexpool;
bool dowatercycles = true, dowindcycles = true;

int main(int argc, char** argv) {
  if (argc < 7) { printf("usage: facade_loop file.soil dim nwater nwind frames out.bin\n"); return 2; }
  try {
    WorldEntry w = loadsoil(argv[1]);
    SCALE = w.scale; SLICE = 2 * SCALE;
    SIZEX = SIZEY = atoi(argv[2]); NWATER = atoi(argv[3]); NWIND = atoi(argv[4]);
    const int frames = atoi(argv[5]);
    srand(SEED);
    WaterParticle::init(SIZEX, SIZEY); WindParticle::init(SIZEX, SIZEY);
    Layermap map(SEED, ivec2(SIZEX, SIZEY), vertexpool, SCALE);
    map.meshpool(vertexpool);
    for (int frame = 0; frame < frames; frame++) {
      // ---- SoilMachine.cpp:287-320, loop body as upstream writes it ----
      if(dowatercycles)
      for(int i = 0; i < NWATER; i++){

        WaterParticle particle(map);

        while(true){
          while(particle.move(map, vertexpool) && particle.interact(map, vertexpool));
          if(!particle.flood(map, vertexpool))
            break;
        }

      }

      if(dowatercycles)
      WaterParticle::seep(map, vertexpool);

      if(dowindcycles)
      for(int i = 0; i < NWIND; i++){
        WindParticle particle(map);
        while(particle.move(map, vertexpool) && particle.interact(map, vertexpool));
      }

      if(dowatercycles){
        WaterParticle::mapfrequency(map);
        float wf = 0.f;
        for (int k = 0; k < SIZEX * SIZEY; k++) wf += WaterParticle::frequency[k];
        WaterParticle::resetfrequency(map);
        printf("frame %d: mean water frequency %.6g\n", frame, wf / (SIZEX * SIZEY));
      }
    }
    map.update(vertexpool);                                   // layermap.h:551-555 through the device mesher
    // dump every column through Layermap::top()/prev (layermap.h:150-152) and the per-cell reads
    FILE* f = fopen(argv[6], "wb");
    if (!f) return 3;
    for (int x = 0; x < SIZEX; x++) for (int y = 0; y < SIZEY; y++) {
      int n = 0;
      for (sec* e = map.top(ivec2(x, y)); e != nullptr; e = e->prev) n++;
      fwrite(&n, sizeof(int), 1, f);
      for (sec* e = map.top(ivec2(x, y)); e != nullptr; e = e->prev) {      // top -> bottom
        const long long t = (long long)e->type;
        fwrite(&t, sizeof(t), 1, f); fwrite(&e->size, 8, 1, f); fwrite(&e->floor, 8, 1, f); fwrite(&e->saturation, 8, 1, f);
      }
      const double h = map.height(ivec2(x, y));               // served by the host mirror after the first reads
      fwrite(&h, 8, 1, f);
    }
    fclose(f);
    const vec3 nb = map.normal(vec2(5.25f, 7.5f));
    printf("facade loop ok: %zu mesh uploads of %zu vertices, normal(5.25,7.5) = %.6f %.6f %.6f\n",
           vertexpool.uploads, vertexpool.nvertices, nb.x, nb.y, nb.z);
  } catch (const Error& e) {
    printf("soilmachine error %d: %s\n", e.code, e.what());
    return e.code == SM_ERR_NOGPU ? 77 : 1;
  } catch (const SoilFileError& e) { printf("%s\n", e.what()); return 2; }
  return 0;
}
