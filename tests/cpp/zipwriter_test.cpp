// Drives bgx_bs::seqset_file_writer from a manifest (tests/test_zip64.py): one member per line,
//   J <name> <hex of the JSON text>      JSON member (create_json -> create_path_contents)
//   A <name> <size>                      array member of <size> zero bytes (create_membuf -> create_path)
//   R <name> <size> <hex marker>         reserved member (a hole), the marker written at its end
#include <fstream>
#include <iostream>
#include <sstream>

#include "bgx_build_seqset.hpp"

static std::string unhex(const std::string& h) {
  std::string o;
  for (size_t i = 0; i + 1 < h.size(); i += 2) o.push_back((char)std::stoi(h.substr(i, 2), nullptr, 16));
  return o;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  bgx_bs::seqset_file_writer w(argv[1]);
  std::ifstream in(argv[2]);
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string kind, name, a, b;
    ls >> kind >> name >> a >> b;
    if (kind == "J") {
      w.add(name, unhex(a));
    } else if (kind == "A") {
      std::string zeros(std::stoull(a), '\0');
      w.add(name, zeros.data(), zeros.size());
    } else if (kind == "R") {
      const uint64_t size = std::stoull(a);
      const uint64_t off = w.reserve(name, size);
      const std::string mark = unhex(b);
      w.write_at(off + size - mark.size(), mark.data(), mark.size());
    }
  }
  std::cout << w.finish() << "\n";
  return 0;
}
