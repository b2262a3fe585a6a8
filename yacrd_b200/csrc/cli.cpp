// cli.cpp — `yacrd`-compatible driver for the detect path (reference src/main.rs:36-137, flags src/cli.rs:39-74).
//
//   yacrd-b200 -i overlaps.paf|.m4|.mhap|report.yacrd -o report.yacrd [-c COVERAGE] [-n NOT_COVERAGE] [-t THREADS]
//
// Same flags, defaults and version string as the reference; the work goes through the C ABI (include/yacrd_b200.h)
// exactly as main.rs drives its trait objects: init (main.rs:57) -> compute_all_bad_part (main.rs:78) -> one report
// line per read (main.rs:80-84) -> the optional editor subcommand (main.rs:87-117):
//   yacrd-b200 -i overlaps.paf -o report.yacrd scrubb|filter|extract|split -i reads.fastq -o edited.fastq
// `-d/--ondisk` selects streamed batches: the reference bounds its working set by flushing the overlap store to disk every
// --ondisk-buffer-size bytes (cli.rs:61-70); here the input goes through the device in chunks of that many bytes of
// intervals (8 bytes each), transfers overlapped with the kernels. The prefix itself is not used: nothing goes to disk.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>

#include "../../include/yacrd_b200.h"

namespace {

void usage(FILE *f) {
    fprintf(f,
            "yacrd %s\n"
            "B200-native detect path of yacrd (chimera / not-covered read detection from all-vs-all overlaps)\n\n"
            "USAGE:\n    yacrd-b200 [OPTIONS] --input <INPUT> --output <OUTPUT>\n\n"
            "OPTIONS:\n"
            "    -i, --input <INPUT>                 overlap file (.paf|.m4|.mhap) or yacrd report (.yacrd)\n"
            "    -o, --output <OUTPUT>               path of the .yacrd report\n"
            "    -c, --coverage <COVERAGE>           if coverage reach this value region is marked as bad [default: 0]\n"
            "    -n, --not-coverage <NOT_COVERAGE>   bad-length / length above which a read is NotCovered [default: 0.8]\n"
            "    -t, --thread <THREADS>              host threads of the parser, 0 = all [default: all]\n"
            "        --read-buffer-size <SIZE>       read buffer of the parser [default: 8192]\n"
            "    -d, --ondisk <PREFIX>               streamed batches: chunks of --ondisk-buffer-size bytes of intervals go\n"
            "                                        through the device, transfers overlapped (nothing is written to disk)\n"
            "        --ondisk-buffer-size <SIZE>     bytes of intervals per chunk with -d [default: 64000000]\n"
            "        --device <N>                    CUDA device ordinal [default: current]\n"
            "        --timing                        phase times on stderr\n"
            "    -h, --help    -V, --version\n\n"
            "SUBCOMMANDS (after the options above):\n"
            "    scrubb  -i <INPUT> -o <OUTPUT>   all bad regions of the reads are removed (fasta|fastq)\n"
            "    filter  -i <INPUT> -o <OUTPUT>   records marked Chimeric or NotCovered are dropped (fasta|fastq|paf|m4)\n"
            "    extract -i <INPUT> -o <OUTPUT>   only those records are kept (fasta|fastq|paf|m4)\n"
            "    split   -i <INPUT> -o <OUTPUT>   Chimeric reads are cut at their interior bad regions (fasta|fastq)\n",
            yb_version());
}

bool parse_u64(const char *s, uint64_t *out) {
    if (!*s) return false;
    char *e = nullptr;
    if (*s == '-') return false;
    const unsigned long long v = strtoull(s, &e, 10);
    if (*e) return false;
    *out = v;
    return true;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, char **argv) {
    std::string input, output, sub_input, sub_output;
    int subcmd = -1;  // yb_editor
    uint64_t coverage = 0, threads = 0, buffer_size = 8192, ondisk_buffer_size = 64000000;
    bool ondisk = false;
    double not_coverage = 0.8;
    int device = -1;
    bool timing = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: The argument '%s' requires a value but none was supplied\n", name);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") {
            usage(stdout);
            return 0;
        } else if (a == "-V" || a == "--version") {
            printf("yacrd %s\n", yb_version());
            return 0;
        } else if (a == "-i" || a == "--input") {
            input = value("--input");
        } else if (a == "-o" || a == "--output") {
            output = value("--output");
        } else if (a == "-c" || a == "--coverage") {
            if (!parse_u64(value("--coverage"), &coverage)) {
                fprintf(stderr, "error: Invalid value for '--coverage <COVERAGE>': invalid digit found in string\n");
                return 2;
            }
        } else if (a == "-n" || a == "--not-coverage") {
            char *e = nullptr;
            const char *v = value("--not-coverage");
            not_coverage = strtod(v, &e);
            if (!*v || *e) {
                fprintf(stderr, "error: Invalid value for '--not-coverage <NOT_COVERAGE>': invalid float literal\n");
                return 2;
            }
        } else if (a == "-t" || a == "--thread") {
            if (!parse_u64(value("--thread"), &threads)) {
                fprintf(stderr, "error: Invalid value for '--thread <THREADS>'\n");
                return 2;
            }
        } else if (a == "--read-buffer-size") {
            if (!parse_u64(value("--read-buffer-size"), &buffer_size)) {
                fprintf(stderr, "error: Invalid value for '--read-buffer-size <BUFFER_SIZE>'\n");
                return 2;
            }
        } else if (a == "-d" || a == "--ondisk") {
            value("--ondisk");
            ondisk = true;
        } else if (a == "--ondisk-buffer-size") {
            if (!parse_u64(value("--ondisk-buffer-size"), &ondisk_buffer_size)) {
                fprintf(stderr, "error: Invalid value for '--ondisk-buffer-size <ONDISK_BUFFER_SIZE>'\n");
                return 2;
            }
        } else if (a == "--device") {
            device = atoi(value("--device"));
        } else if (a == "--timing") {
            timing = true;
        } else if (a == "scrubb" || a == "filter" || a == "extract" || a == "split") {  // cli.rs:77-137: the rest belongs to it
            subcmd = a == "scrubb" ? YB_EDIT_SCRUBB : a == "filter" ? YB_EDIT_FILTER : a == "extract" ? YB_EDIT_EXTRACT : YB_EDIT_SPLIT;
            for (++i; i < argc; ++i) {
                const std::string b = argv[i];
                if (b == "-i" || b == "--input") sub_input = value("--input");
                else if (b == "-o" || b == "--output") sub_output = value("--output");
                else {
                    fprintf(stderr, "error: Found argument '%s' which wasn't expected, or isn't valid in this context\n", b.c_str());
                    return 2;
                }
            }
            if (sub_input.empty() || sub_output.empty()) {
                fprintf(stderr, "error: The following required arguments were not provided:%s%s\n", sub_input.empty() ? "\n    --input <INPUT>" : "",
                        sub_output.empty() ? "\n    --output <OUTPUT>" : "");
                return 2;
            }
        } else {
            fprintf(stderr, "error: Found argument '%s' which wasn't expected, or isn't valid in this context\n", a.c_str());
            return 2;
        }
    }
    if (input.empty() || output.empty()) {
        fprintf(stderr, "error: The following required arguments were not provided:%s%s\n", input.empty() ? "\n    --input <INPUT>" : "",
                output.empty() ? "\n    --output <OUTPUT>" : "");
        return 2;
    }
    const double t0 = now_s();
    yb_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.device = device;
    opts.read_buffer_size = (uint32_t)buffer_size;
    opts.ingest_threads = (uint32_t)threads;
    opts.flags = YB_FLAG_LAZY_DEVICE;  // CUDA starts up on another thread while the input is parsed
    std::thread warm([device] { yb_device_warmup(device); });
    struct Joiner {
        std::thread &t;
        ~Joiner() {
            if (t.joinable()) t.join();
        }
    } joiner{warm};
    yb_ctx *ctx = yb_create(&opts);
    if (!ctx) {
        fprintf(stderr, "Error: %s\n", yb_create_error());
        return 1;
    }
    auto fail = [&](const char *what) {
        fprintf(stderr, "Error: %s\n\nCaused by:\n    %s\n", what, yb_last_error(ctx));
        yb_destroy(ctx);
        return 1;
    };
    if (ondisk) yb_set_chunk_intervals(ctx, (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ondisk_buffer_size / 8, 1024), 0xFFFFFFFFull));
    const double t1 = now_s();
    const bool from_report = yb_file_type(input.c_str()) == 'y';  // main.rs:43-45
    if ((from_report ? yb_init_report(ctx, input.c_str()) : yb_init_file(ctx, input.c_str())) != YB_OK) return fail("reading the input");
    const double t2 = now_s();
    if (yb_compute_all_bad_part(ctx, coverage, not_coverage) != YB_OK) return fail("computing the bad regions");
    const double t3 = now_s();
    if (yb_write_report(ctx, output.c_str()) != YB_OK) return fail("writing the report");
    if (subcmd >= 0 && yb_edit(ctx, subcmd, sub_input.c_str(), sub_output.c_str()) != YB_OK) return fail("running the editor");
    const double t4 = now_s();
    if (timing) {
        yb_stats st;
        yb_get_stats(ctx, &st);
        fprintf(stderr,
                "[yacrd-b200] %llu reads, %llu intervals: context %.3f s | read+parse %.3f s | H2D+kernels+D2H %.3f s | report %.3f s | "
                "NotBad %llu Chimeric %llu NotCovered %llu\n",
                (unsigned long long)st.n_reads, (unsigned long long)st.n_intervals, t1 - t0, t2 - t1, t3 - t2, t4 - t3,
                (unsigned long long)st.n_not_bad, (unsigned long long)st.n_chimeric, (unsigned long long)st.n_not_covered);
    }
    yb_destroy(ctx);
    return 0;
}
