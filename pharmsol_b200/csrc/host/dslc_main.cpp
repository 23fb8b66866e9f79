// dslc — host tool: pharmsol-dsl source -> CUDA C translation unit (used by the build to generate
// the ahead-of-time kernels; the runtime library embeds the same emitter for NVRTC).
#include <fstream>
#include <iostream>
#include <sstream>

#include "dsl.hpp"

int main(int argc, char** argv) {
    bool aot = false, info = false;
    const char* path = nullptr;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--aot") aot = true;
        else if (a == "--info") info = true;
        else path = argv[i];
    }
    if (!path) { std::cerr << "usage: dslc [--aot] [--info] model.pmdsl\n"; return 2; }
    std::ifstream f(path);
    if (!f) { std::cerr << "cannot open " << path << "\n"; return 2; }
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        auto cm = pharmsol::dsl::compile_source(ss.str());
        if (info) { std::cout << cm.model_info_json() << "\n"; return 0; }
        std::vector<std::pair<int, std::string>> entries;
        const int nsolvers = cm.kind == pharmsol::dsl::ModelKind::Ode ? 7 : 1;
        for (int s = 0; s < nsolvers; ++s) entries.emplace_back(s, "psi_entry_" + cm.id + "_s" + std::to_string(s));
        std::cout << cm.cuda_source(entries, aot);
    } catch (const std::exception& e) {
        std::cerr << "dslc: " << path << ": " << e.what() << "\n";
        return 1;
    }
    return 0;
}
