// GL-free stand-ins for the pieces of TinyEngine / vertexpool.h that the reference's
// hot-path headers name in their signatures.  TEST INFRASTRUCTURE (oracle only).
//
//  * Vertex / Vertexpool<Vertex>: every hot-path function takes a Vertexpool<Vertex>&
//    (water.h:43,75; wind.h:54,94; particle.h:24) and Layermap::update writes one Vertex per
//    touched cell (layermap.h:475-549).  The real one is a persistently mapped GL buffer
//    (source/include/vertexpool.h:95-342); here it is a plain std::vector so the write (and the
//    normal() it forces) still happens when the CPU baseline is timed.
//  * SDL_Surface / image:: exist only so io.h:234-252 parses; never called.
#pragma once
#include <vector>
#include <string>
#include <utility>
#include <new>
#include <glm/glm.hpp>

struct Vertex {
  Vertex() {}
  Vertex(glm::vec3 p, glm::vec3 n, glm::vec4 c, int i) {
    position[0] = p.x; position[1] = p.y; position[2] = p.z;
    normal[0] = n.x; normal[1] = n.y; normal[2] = n.z;
    color[0] = c.x; color[1] = c.y; color[2] = c.z; color[3] = c.w;
    index = i;
  }
  float position[3];
  float normal[3];
  float color[4];
  float index;
};

template <typename T> struct Vertexpool {
  std::vector<T> buf;
  std::vector<glm::uint> indices;
  glm::uint base = 0;
  Vertexpool() {}
  Vertexpool(int k, int n) { buf.resize((size_t)k * (size_t)n); }
  glm::uint* section(int, int = 0, glm::vec3 = glm::vec3(0)) { return &base; }
  void unsection(glm::uint*) {}
  void resize(const glm::uint*, int) {}
  void index() {}
  void update() {}
  T* get(glm::uint*, int k) { return &buf[k]; }
  template <typename... A> void fill(glm::uint*, int k, A&&... a) {
    if ((size_t)k < buf.size()) new (&buf[k]) T(std::forward<A>(a)...);
  }
};

struct SDL_Surface;
namespace image {
template <typename F> SDL_Surface* make(F, glm::ivec2) { return nullptr; }
inline void save(SDL_Surface*, std::string) {}
}  // namespace image
