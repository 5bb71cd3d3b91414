// bgx-merge -- `biograph merge` on a B200: several BioGraphs in, one merged BioGraph out.
//
// Takes the flags of MergeSEQSETMain (modules/biograph/biograph_merge.cpp:87-104), runs its pre-flight
// checks with its messages (:106-161) and its stages (:199-400)
//   make_flats -> make_mergemaps -> final_merge -> create_readmaps -> metadata
// through the facade (include/bgx_build_seqset.hpp: seqset_file, seqset_merger -> bgx_merge_seqsets ->
// CUDA; the first three stages are ONE device call, no .flat / .mergemap temp files), and writes
//   <out>/seqset                          merged seqset (every payload member as `biograph merge` writes it)
//   <out>/coverage/<sha1>.readmap         one migrated readmap per input sample (make_readmap::fast_migrate)
//   <out>/metadata/bg_info.json           accession id ("a+b" unless --id), samples, command history
//   <out>/qc/<accession>_create_log.txt, <accession>_kmer_quality_report.html   copied from the inputs
//   <out>/qc/merge_stats.json, merge_log.txt
// There is no CPU fallback: the merge needs a CUDA device.
#include <dirent.h>
#include <sys/stat.h>

#include <map>
#include <set>
#include <sstream>

#include "bgx_build_seqset.hpp"
#include "cli_util.hpp"

using namespace bgx_cli;

namespace {

const char* kVersion = "7.1.2-dev";  // versions.bzl:12 BIOGRAPH_VERSION of the reference commit

// ---- just enough JSON to read metadata/bg_info.json (biograph_metadata, biograph_dir.h:17-33) ----------------
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json* get(const std::string& k) const {
    for (const auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

struct JsonParser {
  const std::string& s;
  size_t p = 0;
  explicit JsonParser(const std::string& text) : s(text) {}
  void ws() { while (p < s.size() && isspace((unsigned char)s[p])) ++p; }
  [[noreturn]] void bad() { throw std::runtime_error("bad JSON"); }
  std::string string() {
    if (s[p] != '"') bad();
    std::string o;
    for (++p; p < s.size() && s[p] != '"'; ++p) {
      if (s[p] != '\\') { o += s[p]; continue; }
      if (++p >= s.size()) bad();
      switch (s[p]) {
        case 'n': o += '\n'; break;
        case 't': o += '\t'; break;
        case 'r': o += '\r'; break;
        case 'b': o += '\b'; break;
        case 'f': o += '\f'; break;
        case 'u': {
          if (p + 4 >= s.size()) bad();
          const unsigned c = (unsigned)strtoul(s.substr(p + 1, 4).c_str(), nullptr, 16);
          p += 4;
          if (c < 0x80) o += (char)c;
          else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 0x3F)); }
          else { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
          break;
        }
        default: o += s[p];
      }
    }
    if (p >= s.size()) bad();
    ++p;
    return o;
  }
  Json value() {
    ws();
    if (p >= s.size()) bad();
    Json j;
    if (s[p] == '{') {
      j.kind = Json::Obj;
      ++p; ws();
      if (s[p] == '}') { ++p; return j; }
      for (;;) {
        ws();
        std::string k = string();
        ws();
        if (s[p++] != ':') bad();
        j.obj.emplace_back(k, value());
        ws();
        if (s[p] == ',') { ++p; continue; }
        if (s[p] == '}') { ++p; return j; }
        bad();
      }
    }
    if (s[p] == '[') {
      j.kind = Json::Arr;
      ++p; ws();
      if (s[p] == ']') { ++p; return j; }
      for (;;) {
        j.arr.push_back(value());
        ws();
        if (s[p] == ',') { ++p; continue; }
        if (s[p] == ']') { ++p; return j; }
        bad();
      }
    }
    if (s[p] == '"') { j.kind = Json::Str; j.str = string(); return j; }
    const size_t b = p;
    while (p < s.size() && (isalnum((unsigned char)s[p]) || s[p] == '-' || s[p] == '+' || s[p] == '.')) ++p;
    if (b == p) bad();
    j.str = s.substr(b, p - b);
    j.kind = j.str == "null" ? Json::Null : (j.str == "true" || j.str == "false") ? Json::Bool : Json::Num;
    return j;
  }
};

std::string slurp(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  std::stringstream ss;
  ss << in.rdbuf();
  return ss.str();
}

bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

void copy_file(const std::string& from, const std::string& to) {
  std::ifstream in(from, std::ios::binary);
  std::ofstream out(to, std::ios::binary);
  out << in.rdbuf();
  if (!out.good()) throw std::runtime_error("cannot copy " + from + " to " + to);
}

// biograph_dir(path, READ_BGDIR) (modules/bio_base/biograph_dir.cpp:13-23,39-55,87-101)
struct BgDir {
  std::string path, version, biograph_id, accession_id;
  std::vector<std::pair<std::string, std::string>> samples;  // accession -> readmap sha1, in key order (std::map)
  std::vector<std::string> command_history;
  bool valid = false;
  explicit BgDir(std::string p) : path(std::move(p)) {
    while (path.size() > 1 && path.back() == '/') path.pop_back();
    bool ok = is_dir(path);
    for (const char* d : {"metadata", "coverage", "qc"}) ok = ok && exists(path + "/" + d);
    const std::string meta = path + "/metadata/bg_info.json";
    if (ok && exists(meta)) {
      try {
        const std::string text = slurp(meta);
        JsonParser jp(text);
        const Json j = jp.value();
        if (const Json* v = j.get("version")) version = v->str;
        if (const Json* v = j.get("biograph_id")) biograph_id = v->str;
        if (const Json* v = j.get("accession_id")) accession_id = v->str;
        std::map<std::string, std::string> sm;
        if (const Json* v = j.get("samples")) for (const auto& kv : v->obj) sm[kv.first] = kv.second.str;
        samples.assign(sm.begin(), sm.end());
        if (const Json* v = j.get("command_history")) for (const Json& c : v->arr) command_history.push_back(c.str);
      } catch (...) {
        throw std::runtime_error("Could not parse biograph metadata: " + meta);
      }
      valid = true;
    }
    if (!valid) throw std::runtime_error("Attempted to open " + path + " but the BioGraph was not valid. Cannot continue.");
  }
  std::string seqset() const { return path + "/seqset"; }
  std::string readmap(const std::string& rm) const { return path + "/coverage/" + rm + ".readmap"; }
};

struct Args {
  std::string out, id, stats_file;
  std::vector<std::string> in;
  bool force = false;
  int device = 0;
  uint64_t parallel_splits = 0;
  bool list_inputs = false;   // test hook: pre-flight + input reading only, no GPU
};

void usage() {
  std::cerr << "bgx-merge version " << kVersion << " (" << bgx_version() << ")\n\n"
            << "Usage: bgx-merge [OPTIONS] --out <merged biograph> --in <source biograph> <source biograph> [...]\n\n"
               "Merge BioGraphs. Produces a single merged BioGraph with coverage data for every\n"
               "sample in the input BioGraphs.\n\n"
               "  --out arg            Output merged BioGraph\n"
               "  --in arg             Input biographs to merge\n"
               "  --id arg             Optional accession ID for the merged BioGraph\n"
               "  -f, --force          Overwrite existing BioGraph\n"
               "  --device arg (=0)    CUDA device ordinal\n"
               "  --parallel-splits arg (=100000)  chunking of seqset_merger's prev-bit pass (the reference's\n"
               "                       g_parallel_splits; 1 = the placement `biograph create` uses)\n"
               "  (accepted for compatibility, no effect here: --tmp, --keep-tmp, --threads, --max-mem, --cache, --stats)\n";
}

Args parse(int argc, char** argv) {
  Args a;
  std::vector<std::string> positional;
  bool in_multi = false;  // --in is multitoken (biograph_merge.cpp:90)
  auto need = [&](int& i) -> std::string {
    if (i + 1 >= argc) die(std::string("the required argument for option '") + argv[i] + "' is missing");
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i) {
    std::string o = argv[i], v;
    const size_t eq = o.find('=');
    bool has_v = false;
    if (o.rfind("--", 0) == 0 && eq != std::string::npos) { v = o.substr(eq + 1); o = o.substr(0, eq); has_v = true; }
    auto val = [&]() { return has_v ? v : need(i); };
    if (o.rfind("-", 0) != 0) {
      if (in_multi) a.in.push_back(o); else positional.push_back(o);
      continue;
    }
    in_multi = false;
    if (o == "--out") a.out = val();
    else if (o == "--in") { a.in.push_back(val()); in_multi = true; }
    else if (o == "--id") a.id = val();
    else if (o == "--force" || o == "-f") a.force = true;
    else if (o == "--device") a.device = atoi(val().c_str());
    else if (o == "--parallel-splits") a.parallel_splits = strtoull(val().c_str(), nullptr, 10);
    else if (o == "--list-inputs") a.list_inputs = true;
    else if (o == "--stats") a.stats_file = val();
    else if (o == "--tmp" || o == "--threads" || o == "--max-mem") (void)val();
    else if (o == "--keep-tmp" || o == "--cache" || o == "--debug") {}
    else if (o == "--help" || o == "-h") { usage(); exit(0); }
    else die("unrecognised option '" + o + "'");
  }
  // positional: out, then the inputs (biograph_merge.cpp:100-101)
  size_t pi = 0;
  if (a.out.empty() && pi < positional.size()) a.out = positional[pi++];
  for (; pi < positional.size(); ++pi) a.in.push_back(positional[pi]);
  if (a.out.empty()) die("the option '--out' is required but missing");
  if (a.in.empty()) die("the option '--in' is required but missing");
  return a;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    Args a = parse(argc, argv);
    // ---- pre-flight (biograph_merge.cpp:106-161) ----------------------------------------------------------------
    std::set<std::string> in_bg_ids, in_accessions, in_sample_ids;
    std::vector<BgDir> in_dirs;
    bool use_full_ids = false;
    for (const std::string& in_file : a.in) {
      // the constructor throws for an invalid directory, so :111-113's own message is never reached
      std::unique_ptr<BgDir> bg(new BgDir(in_file));
      if (in_bg_ids.count(bg->biograph_id)) { std::cerr << "Duplicate BioGraph ID for '" << in_file << "', skipping.\n"; continue; }
      if (in_accessions.count(bg->accession_id)) {
        std::cerr << "Duplicate Accession ID '" << bg->accession_id << "' for '" << in_file << "', skipping.\n";
        continue;
      }
      if (bg->samples.empty()) throw std::runtime_error("No sample metadata found for '" + in_file + "'. Cannot continue.");
      in_bg_ids.insert(bg->biograph_id);
      in_accessions.insert(bg->accession_id);
      if (!use_full_ids)
        for (const auto& smp : bg->samples) {
          if (in_sample_ids.count(smp.first)) { use_full_ids = true; break; }
          in_sample_ids.insert(smp.first);
        }
      in_dirs.push_back(*bg);
    }
    if (in_dirs.size() < 2) throw std::runtime_error("Merge requires two or more unique BioGraphs.");
    if (a.list_inputs) {
      // test hook: what the merge would read, one JSON line per input (no GPU, nothing written)
      for (const BgDir& d : in_dirs) {
        std::cout << "{\"path\":" << json_str(d.path) << ",\"biograph_id\":" << json_str(d.biograph_id) << ",\"accession_id\":" << json_str(d.accession_id)
                  << ",\"samples\":{";
        for (size_t i = 0; i < d.samples.size(); ++i)
          std::cout << (i ? "," : "") << json_str(d.samples[i].first) << ":" << json_str(d.samples[i].second);
        std::cout << "},\"command_history\":" << d.command_history.size();
        struct stat st;
        if (stat(d.seqset().c_str(), &st) == 0) {
          bgx_bs::seqset_file f(d.seqset());
          std::cout << ",\"entries\":" << f.size() << ",\"max_read_len\":" << f.max_read_len() << ",\"uuid\":" << json_str(f.uuid());
        }
        std::cout << ",\"use_full_ids\":" << (use_full_ids ? "true" : "false") << "}\n";
      }
      return 0;
    }
    if (!a.force && exists(a.out)) {
      std::cerr << "Refusing to overwrite '" + a.out + "'. Use --force to override.\n";
      return 1;
    }
    // biograph_dir(m_out, CREATE_BGDIR)
    mkdir(a.out.c_str(), 0777);
    for (const char* d : {"metadata", "coverage", "qc", "analysis"}) mkdir((a.out + "/" + d).c_str(), 0777);
    if (!is_dir(a.out + "/qc"))
      throw std::runtime_error("Attempted to create " + a.out + " but the resulting biograph was not valid. Cannot continue.");
    if (a.stats_file.empty()) a.stats_file = a.out + "/qc/merge_stats.json";
    std::ofstream log(a.out + "/qc/merge_log.txt");
    auto splog = [&](const std::string& m) {
      const time_t now = time(nullptr);
      char ts[32];
      strftime(ts, sizeof(ts), "%Y-%m-%d %H:%M:%S", localtime(&now));
      log << ts << " " << m << "\n";
      log.flush();
    };
    {
      std::string cmd;
      for (int i = 0; i < argc; ++i) cmd += std::string(i ? " " : "") + argv[i];
      splog("bgx-merge " + std::string(kVersion) + " (" + bgx_version() + "): " + cmd);
    }
    Stages stages;
    const auto t_total = std::chrono::steady_clock::now();

    // ---- make_flats + make_mergemaps + final_merge (:199-290): one device call ---------------------------------------
    stages.start();
    bgx_bs::count_kmer_options ko;
    ko.device = a.device;
    bgx_bs::session sess(ko);
    std::vector<std::unique_ptr<bgx_bs::seqset_file>> seqsets;
    std::vector<const bgx_bs::seqset_file*> ptrs;
    for (const BgDir& d : in_dirs) {
      std::cerr << d.path << std::endl;
      splog("Building flat seqset for " + d.path);
      seqsets.emplace_back(new bgx_bs::seqset_file(d.seqset()));
      ptrs.push_back(seqsets.back().get());
    }
    stages.end("make_flats");
    stages.start();
    std::cerr << "Creating merge maps" << std::endl;
    bgx_bs::seqset_merger merger(sess, ptrs, a.parallel_splits);
    merger.build();
    splog(fmt("%lu entries in resultant merge; writing mergemaps", (unsigned long)merger.total_merged_entries()));
    stages.end("make_mergemaps");
    stages.start();
    std::cerr << "Generating merged BioGraph" << std::endl;
    const std::string uuid = make_uuid();
    const bgx_bs::seqset_tables tables = merger.write_seqset(a.out + "/seqset", uuid);
    splog(fmt("Creating merged seqset with %lu entries, %u maximum entry length", (unsigned long)tables.num_entries, tables.max_entry_len));
    stages.end("final_merge");

    // ---- create_readmaps (:292-330) ---------------------------------------------------------------------------------------
    stages.start();
    std::map<std::string, std::string> samples;
    for (size_t i = 0; i < in_dirs.size(); ++i) {
      const BgDir& d = in_dirs[i];
      for (const auto& smp : d.samples) {
        splog("Migrating " + d.biograph_id + ":" + smp.second);
        std::cerr << "Coverage: " << d.accession_id << " (" << smp.first << ")" << std::endl;
        const std::string tmp = a.out + "/coverage/tmp.readmap";
        remove(tmp.c_str());
        merger.fast_migrate((unsigned)i, d.readmap(smp.second), tmp, uuid);
        const std::string sha = sha1_file(tmp);
        if (rename(tmp.c_str(), (a.out + "/coverage/" + sha + ".readmap").c_str()) != 0) throw std::runtime_error("cannot rename the readmap");
        samples[use_full_ids ? d.accession_id + ":" + smp.first : smp.first] = sha;
      }
    }
    stages.end("create_readmaps");

    // ---- metadata (:332-400) ----------------------------------------------------------------------------------------------------
    stages.start();
    std::string accession = a.id;
    if (accession.empty())
      for (const BgDir& d : in_dirs) accession += (accession.empty() ? "" : "+") + d.accession_id;
    std::vector<std::string> history;
    for (size_t i = 0; i < in_dirs.size(); ++i) {
      const BgDir& d = in_dirs[i];
      // file_info().command_line_str(): the command line that wrote the input's seqset, space-joined
      {
        std::string cmd;
        try {
          const std::string fi = seqsets[i]->file_info();
          JsonParser jp(fi);
          const Json j = jp.value();
          if (const Json* v = j.get("command_line"))
            for (const Json& c : v->arr) cmd += (cmd.empty() ? "" : " ") + c.str;
        } catch (...) {
        }
        history.push_back(cmd);
      }
      for (const std::string& c : d.command_history) history.push_back(c);
      const std::string qc = d.path + "/qc";
      if (exists(qc + "/create_log.txt")) copy_file(qc + "/create_log.txt", a.out + "/qc/" + d.accession_id + "_create_log.txt");
      if (exists(qc + "/kmer_quality_report.html"))
        copy_file(qc + "/kmer_quality_report.html", a.out + "/qc/" + d.accession_id + "_kmer_quality_report.html");
      if (DIR* dir = opendir(qc.c_str())) {
        while (dirent* e = readdir(dir)) {
          const std::string f = e->d_name;
          if (f.find("_log.txt") != std::string::npos || f.find("_kmer_quality_report.html") != std::string::npos) {
            const std::string dest = a.out + "/qc/" + d.accession_id + "_" + f;
            if (!exists(dest)) copy_file(qc + "/" + f, dest);
          }
        }
        closedir(dir);
      }
    }
    {
      std::ofstream os(a.out + "/metadata/bg_info.json");
      os << "{\"accession_id\":" << json_str(accession) << ",\"biograph_id\":" << json_str(uuid) << ",\"command_history\":[";
      for (size_t i = 0; i < history.size(); ++i) os << (i ? "," : "") << json_str(history[i]);
      os << "],\"samples\":{";
      size_t k = 0;
      for (const auto& smp : samples) os << (k++ ? "," : "") << json_str(smp.first) << ":" << json_str(smp.second);
      os << "},\"version\":" << json_str(kVersion) << "}";
      if (!os.good()) throw std::runtime_error("Could not write to " + a.out + "/metadata/bg_info.json");
    }
    stages.end("metadata");
    {
      std::ofstream os(a.stats_file);
      os << "{\"command\":\"merge\",\"version\":" << json_str(kVersion) << ",\"accession_id\":" << json_str(accession) << ",\"samples\":"
         << samples.size() << ",\"uuid\":" << json_str(uuid) << ",\"entries\":" << tables.num_entries << ",\"timings\":[";
      for (const auto& s : stages.t) os << "{" << json_str(s.first) << ":" << (long)s.second << "},";
      os << "{\"total\":" << (long)std::chrono::duration<double>(std::chrono::steady_clock::now() - t_total).count() << "}]}";
    }
    std::cerr << std::endl << a.out << " created." << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
}
