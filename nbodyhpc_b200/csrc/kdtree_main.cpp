// kdtree_main -- the reference's CLI benchmark (kdtree/src/cpp/main.cpp:125-175) over the B200
// drop-in: same flags, same input (Philox points of seed 42, or a raw float32 xyz file), same
// report lines.  The reference queries the first `num-queries` points one by one on one thread
// (main.cpp:51-93); here they are answered as one batch on the GPU, and the visited-points
// statistic comes from the on-device replay of the reference's traversal (nbk_tree_stats).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <kdtree/kdtree.hpp>
#include <kdtree/kdtree_utils.hpp>

namespace {

struct Options {
    uint32_t num_points = 10000000; // main.cpp:130-137 defaults
    int num_neighbors = 16;
    uint32_t num_queries = 500000;
    int threads = -1;
    int leaf_size = 64;
    bool periodic = false;
    float box_size = 1.0f;
    std::string file;
};

void usage() {
    std::cout << "KD-tree benchmark (B200)\nUsage:\n  kdtree [OPTION...]\n\n"
                 "  -n, --num-points arg     Number of points to use (default: 10000000)\n"
                 "      --num-neighbors arg  Number of neighbors to find (default: 16)\n"
                 "  -q, --num-queries arg    Number of queries to perform (default: 500000)\n"
                 "  -t, --threads arg        Accepted for compatibility; the GPU does the work (default: -1)\n"
                 "      --leaf-size arg      Size of kd-tree leaves (default: 64)\n"
                 "      --periodic           Use periodic boundary conditions\n"
                 "      --box_size arg       Box size when using periodic boundary conditions (default: 1.0)\n"
                 "  -f, --file arg           Raw float32 xyz file to use for benchmarking\n"
                 "  -h, --help               Print help\n";
}

bool parse(int argc, char **argv, Options &o) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], value;
        const size_t eq = a.find('=');
        bool has_value = false;
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) {
            value = a.substr(eq + 1);
            a = a.substr(0, eq);
            has_value = true;
        }
        auto next = [&]() -> std::string {
            if (has_value) return value;
            if (i + 1 >= argc) throw std::runtime_error("missing value for " + a);
            return argv[++i];
        };
        if (a == "-h" || a == "--help") return false;
        else if (a == "-n" || a == "--num-points") o.num_points = (uint32_t)std::stoul(next());
        else if (a == "--num-neighbors") o.num_neighbors = std::stoi(next());
        else if (a == "-q" || a == "--num-queries") o.num_queries = (uint32_t)std::stoul(next());
        else if (a == "-t" || a == "--threads") o.threads = std::stoi(next());
        else if (a == "--leaf-size") o.leaf_size = std::stoi(next());
        else if (a == "--periodic") o.periodic = has_value ? (value != "false" && value != "0") : true;
        else if (a == "--box_size") o.box_size = std::stof(next());
        else if (a == "-f" || a == "--file") o.file = next();
        else throw std::runtime_error("unknown option " + a);
    }
    return true;
}

std::vector<std::array<float, 3>> read_array_from_file(std::string const &path) { // main.cpp:101-112
    std::ifstream file(path, std::ios::in | std::ios::binary | std::ios::ate);
    if (!file) throw std::runtime_error("cannot open " + path);
    const auto bytes = file.tellg();
    std::vector<std::array<float, 3>> positions(static_cast<size_t>(bytes) / (sizeof(float) * 3));
    file.seekg(0, std::ios::beg);
    file.read(reinterpret_cast<char *>(positions.data()), positions.size() * sizeof(float) * 3);
    return positions;
}

} // namespace

int main(int argc, char **argv) {
    using namespace wenda::kdtree;
    Options o;
    try {
        if (!parse(argc, argv, o)) {
            usage();
            return 0;
        }
        std::vector<std::array<float, 3>> positions;
        if (o.file.empty()) {
            std::cout << "Benchmarking kdtree with " << o.num_points << " points" << std::endl;
            positions = fill_random_positions(o.num_points, 42); // main.cpp:96
        } else {
            std::cout << "Benchmarking kdtree with data from: " << o.file << std::endl;
            positions = read_array_from_file(o.file);
        }
        std::chrono::high_resolution_clock clock;
        KDTreeConfiguration config{.leaf_size = o.leaf_size, .max_threads = o.threads};
        const auto b0 = clock.now();
        KDTree tree(positions, config);
        const auto b1 = clock.now();

        const size_t nq = std::min<size_t>(o.num_queries, positions.size());
        const size_t k = static_cast<size_t>(o.num_neighbors);
        std::vector<float> dist(nq * k);
        std::vector<uint32_t> idx(nq * k);
        tcb::span<const std::array<float, 3>> queries(positions.data(), nq);
        const auto q0 = clock.now();
        if (o.periodic) tree.find_closest_batch(queries, k, dist.data(), idx.data(), L2PeriodicDistance<float>{o.box_size});
        else tree.find_closest_batch(queries, k, dist.data(), idx.data(), L2Distance{});
        const auto q1 = clock.now();

        float total_distance = 0; // self-query: the nearest neighbour is the point itself (main.cpp:78,84-86)
        for (size_t i = 0; i < nq; ++i) total_distance += dist[i * k];
        if (total_distance != 0)
            std::cout << "Total distance was not 0! Got instead: " << total_distance << std::endl;

        KDTreeQueryStatistics stats;
        std::vector<float> d2(nq * k);
        std::vector<uint32_t> i2(nq * k);
        if (o.periodic) tree.find_closest_batch(queries, k, d2.data(), i2.data(), L2PeriodicDistance<float>{o.box_size}, &stats);
        else tree.find_closest_batch(queries, k, d2.data(), i2.data(), L2Distance{}, &stats);

        const std::chrono::duration<double> build = b1 - b0, query = q1 - q0;
        std::cout << "Build time: " << build.count() << "s" << std::endl;
        std::cout << "Query time: " << query.count() << "s" << std::endl;
        std::cout << "Query performance: " << nq / query.count() << " qps" << std::endl;
        std::cout << "Points visited proportion: "
                  << static_cast<double>(stats.points_visited) / (static_cast<double>(positions.size()) * nq) * 100
                  << "%" << std::endl;
    } catch (std::exception const &e) {
        std::cerr << "kdtree_main: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
