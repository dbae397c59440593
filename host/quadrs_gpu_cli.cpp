// quadrs_gpu -- command-line front end with the reference's grammar, running the DSP chain on a B200.
//
// Mirrors src/bin/quadrs.rs (usage text, the fold over commands) and src/args.rs (flag grammar, SI
// suffixes, filename sniffing) of FauxFaux/quadrs, so the README command lines work unchanged:
//
//   quadrs_gpu from fsk-example.sr21M.fc32 shift 280000 lowpass -power 200 -decimate 32 200000
//              sparkfft -width 64 -stride 16
//
// All sample arithmetic happens in libquadrs_gpu.so (include/quadrs_gpu.h).  `ui` / `eui` are the
// reference's GUIs and are not provided.  `--parse-only` (first argument) prints the parsed commands
// instead of executing them, so the grammar can be tested without a GPU.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <optional>
#include <regex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/quadrs_gpu.h"

namespace {

struct Err : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum class Op { From, Shift, LowPass, SparkFft, Bucket, Write, Gen, Ui, Eui };

struct Command { // Operation, src/lib.rs:25-59
    Op op;
    std::string filename;
    int format = 0;
    uint64_t sample_rate = 0;
    int64_t frequency = 0;
    uint64_t size = 40, decimate = 8;
    uint64_t width = 128, stride = 128;
    std::optional<float> min, max;
    uint64_t levels = 0;
    bool overwrite = false;
    std::string prefix;
    double seconds = 1.0;
    std::vector<int64_t> cos;
};

// usage(), src/bin/quadrs.rs:9-28
void usage(const char *us)
{
    printf("usage: %s \\\n", us);
    printf("    from [-sr SAMPLE_RATE] [-format cf32|cs8|cu8|cs16] FILENAME.sr32k.cf32 \\\n");
    printf("   shift [-]FREQUENCY \\\n");
    printf(" lowpass [-power 20] [-decimate 8] FREQUENCY \\\n");
    printf("sparkfft [-width 128] [-stride =width] [-range LOW:HIGH] \\\n");
    printf("  bucket [-width 128] [-stride =width] [-by freq] COUNT \\\n");
    printf("   write [-overwrite no] FILENAME_PREFIX \\\n");
    printf("     gen [-cos FREQUENCY]* [-len 1 (second)] SAMPLE_RATE \\\n");
    printf("\n\nFormats:\n\n");
    printf(" * cf32: complex (little endian) floats, 32-bit (GNU-Radio, gqrx)\n");
    printf(" *  cs8: complex      signed (integers),  8-bit (HackRF)\n");
    printf(" *  cu8: complex    unsigned (integers),  8-bit (RTL-SDR)\n");
    printf(" * cs16: complex      signed (integers), 16-bit (Fancy)\n\n");
}

// find_multiplication_suffix, src/args.rs:335-352
std::pair<std::string, uint64_t> split_suffix(const std::string &s)
{
    if (s.empty()) return {s, 1};
    switch (s.back()) {
    case 'k': return {s.substr(0, s.size() - 1), 1000};
    case 'M': return {s.substr(0, s.size() - 1), 1000000};
    case 'G': return {s.substr(0, s.size() - 1), 1000000000};
    default: return {s, 1};
    }
}

uint64_t parse_si_u64(const std::string &s) // src/args.rs:365-372
{
    auto [v, mul] = split_suffix(s);
    if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos) {
        if (!v.empty() && v[0] == '+' && v.find_first_not_of("0123456789", 1) == std::string::npos && v.size() > 1) v = v.substr(1);
        else throw Err("invalid digit found in string");
    }
    errno = 0;
    char *end = nullptr;
    const unsigned long long p = strtoull(v.c_str(), &end, 10);
    if (errno || *end) throw Err("number too large to fit in target type");
    if (mul != 1 && p > UINT64_MAX / mul) throw Err("unit is out of range: " + s);
    return p * mul;
}

int64_t parse_si_i64(const std::string &s) // src/args.rs:354-363
{
    auto [v, mul] = split_suffix(s);
    size_t i = (!v.empty() && (v[0] == '-' || v[0] == '+')) ? 1 : 0;
    if (v.size() == i || v.find_first_not_of("0123456789", i) != std::string::npos) throw Err("invalid digit found in string");
    errno = 0;
    const long long p = strtoll(v.c_str(), nullptr, 10);
    if (errno) throw Err("number too large to fit in target type");
    long long out;
    if (__builtin_mul_overflow(p, static_cast<long long>(mul), &out)) throw Err("unit is out of range: " + s);
    return out;
}

double parse_si_f64(const std::string &s) // src/args.rs:374-380
{
    auto [v, mul] = split_suffix(s);
    char *end = nullptr;
    const double p = strtod(v.c_str(), &end);
    if (v.empty() || *end) throw Err("invalid float literal");
    return p * static_cast<double>(mul);
}

bool parse_bool(const std::string &s) // src/args.rs:382-390
{
    if (s == "true" || s == "yes" || s == "y") return true;
    if (s == "false" || s == "no" || s == "n") return false;
    throw Err("unacceptable boolean value: '" + s + "'");
}

std::optional<int> format_from_ext(const std::string &e) // guess_from_extension, src/args.rs:392-402
{
    if (e == "cf32" || e == "fc32") return QD_FMT_CF32;
    if (e == "cs8" || e == "sc8" || e == "c8") return QD_FMT_CS8;
    if (e == "cu8" || e == "su8") return QD_FMT_CU8;
    if (e == "cs16" || e == "sc16" || e == "c16") return QD_FMT_CS16;
    return std::nullopt;
}

// guess_details / guess_format_from_name, src/args.rs:65-135,328-333
void guess_details(const std::string &filename, const std::optional<std::string> &sr_override,
                   const std::optional<std::string> &fmt_override, uint64_t *rate, int *format)
{
    std::optional<std::string> sample_rate;
    std::optional<int> fmt;
    std::smatch m;
    static const std::regex sr_re(R"(\bsr([0-9]+[kMG]?)\b)");
    if (std::regex_search(filename, m, sr_re)) sample_rate = m[1].str();
    static const std::regex gqrx_re(R"(gqrx_.*?_[0-9]+_([0-9]+)_fc.raw)");
    if (std::regex_search(filename, m, gqrx_re)) {
        sample_rate = m[1].str();
        fmt = QD_FMT_CF32;
    }
    static const std::regex rtl_re(R"(g\d+_\d+(?:\.\d+)?M_(\d+k).cu8)");
    if (std::regex_search(filename, m, rtl_re)) {
        sample_rate = m[1].str();
        fmt = QD_FMT_CU8;
    }
    const size_t dot = filename.rfind('.');
    if (dot != std::string::npos)
        if (auto g = format_from_ext(filename.substr(dot + 1))) fmt = g;
    if (sr_override) sample_rate = sr_override;
    if (fmt_override) {
        fmt = format_from_ext(*fmt_override);
        if (!fmt) throw Err("unrecognised extension: \"" + *fmt_override + "\"");
    }
    if (!sample_rate) throw Err("unable to guess sample rate from filename \"" + filename + "\", please specify it");
    *rate = parse_si_u64(*sample_rate);
    if (!fmt) throw Err("unable to guess format from filename \"" + filename + "\", please specify it");
    *format = *fmt;
}

using Flags = std::map<std::string, std::vector<std::string>>;

// read_just_args, src/args.rs:404-445: `-name value` pairs until a token that is not a flag; a token
// whose third character is a digit is a negative number, not a flag
Flags read_just_args(const std::vector<std::string> &a, size_t *i)
{
    Flags ret;
    while (*i < a.size()) {
        const std::string &opt = a[*i];
        if (opt.empty() || opt[0] != '-') break;
        if (opt.size() > 2 && isdigit(static_cast<unsigned char>(opt[2]))) break;
        ++*i;
        if (*i >= a.size()) throw Err(opt + " requires an argument");
        if (a[*i].empty()) throw Err(opt + " requires a non-empty argument");
        ret[opt.substr(1)].push_back(a[*i]);
        ++*i;
    }
    return ret;
}

std::map<std::string, std::string> no_duplicates(const Flags &f) // src/args.rs:447-454
{
    std::map<std::string, std::string> r;
    for (auto &kv : f) {
        if (kv.second.size() != 1) throw Err("'-" + kv.first + "' specified more than once");
        r[kv.first] = kv.second[0];
    }
    return r;
}

std::optional<std::string> take(std::map<std::string, std::string> &m, const char *k)
{
    auto it = m.find(k);
    if (it == m.end()) return std::nullopt;
    std::string v = it->second;
    m.erase(it);
    return v;
}

void ensure_empty(const std::map<std::string, std::string> &m)
{
    if (m.empty()) return;
    std::string keys;
    for (auto &kv : m) keys += (keys.empty() ? "\"" : ", \"") + kv.first + "\"";
    throw Err("invalid flags: [" + keys + "]");
}

std::string next_arg(const std::vector<std::string> &a, size_t *i, const char *what)
{
    if (*i >= a.size()) throw Err(what);
    return a[(*i)++];
}

// args::parse, src/args.rs:19-45
std::vector<Command> parse(const std::vector<std::string> &a)
{
    std::vector<Command> out;
    size_t i = 0;
    while (i < a.size()) {
        const std::string cmd = a[i++];
        Command c{};
        try {
            Flags raw = read_just_args(a, &i);
            if (cmd == "from") { // parse_from, :47-63
                auto m = no_duplicates(raw);
                c.op = Op::From;
                c.filename = next_arg(a, &i, "'from' requires a filename argument");
                auto sr = take(m, "sr"), fmt = take(m, "format");
                ensure_empty(m);
                guess_details(c.filename, sr, fmt, &c.sample_rate, &c.format);
            } else if (cmd == "shift") { // parse_shift, :137-149
                if (!no_duplicates(raw).empty()) throw Err("'shift' has no named arguments");
                c.op = Op::Shift;
                c.frequency = parse_si_i64(next_arg(a, &i, "'shift' requires a frequency argument"));
            } else if (cmd == "lowpass") { // parse_lowpass, :151-184
                auto m = no_duplicates(raw);
                c.op = Op::LowPass;
                c.frequency = static_cast<int64_t>(parse_si_u64(next_arg(a, &i, "'lowpass' requires a frequency argument")));
                if (auto p = take(m, "power")) {
                    const uint64_t v = parse_si_u64(*p);
                    if (v > UINT64_MAX / 2) throw Err("power is too large");
                    c.size = v * 2;
                } else {
                    c.size = 40;
                }
                c.decimate = 8;
                if (auto d = take(m, "decimate")) c.decimate = parse_si_u64(*d);
                ensure_empty(m);
            } else if (cmd == "sparkfft") { // parse_sparkfft, :186-226
                auto m = no_duplicates(raw);
                c.op = Op::SparkFft;
                c.width = 128;
                if (auto w = take(m, "width")) c.width = parse_si_u64(*w);
                c.stride = c.width;
                if (auto s = take(m, "stride")) c.stride = parse_si_u64(*s);
                if (auto r = take(m, "range")) {
                    const size_t colon = r->find(':');
                    if (colon == std::string::npos) throw Err("range argument must contain a ':': '" + *r + "'");
                    char *end = nullptr;
                    const std::string lo = r->substr(0, colon), hi = r->substr(colon + 1);
                    c.min = strtof(lo.c_str(), &end);
                    if (lo.empty() || *end) throw Err("invalid float literal");
                    c.max = strtof(hi.c_str(), &end);
                    if (hi.empty() || *end) throw Err("invalid float literal");
                }
                ensure_empty(m);
            } else if (cmd == "bucket") { // parse_bucket, :228-263
                auto m = no_duplicates(raw);
                c.op = Op::Bucket;
                c.levels = parse_si_u64(next_arg(a, &i, "bucket usage: bucket -by freq [number-of-buckets]"));
                c.width = 128;
                if (auto w = take(m, "width")) c.width = parse_si_u64(*w);
                c.stride = c.width;
                if (auto s = take(m, "stride")) c.stride = parse_si_u64(*s);
                auto by = take(m, "by");
                if (!by || *by != "freq") throw Err("must bucket -by freq, not " + (by ? "Some(\"" + *by + "\")" : std::string("None")));
                ensure_empty(m);
            } else if (cmd == "write") { // parse_write, :265-283
                auto m = no_duplicates(raw);
                c.op = Op::Write;
                if (auto o = take(m, "overwrite")) c.overwrite = parse_bool(*o);
                ensure_empty(m);
                c.prefix = next_arg(a, &i, "'lowpass' requires a frequency argument"); // (sic) args.rs:279
            } else if (cmd == "gen") { // parse_gen, :285-323: -cos may repeat
                c.op = Op::Gen;
                auto it = raw.find("cos");
                if (it == raw.end()) throw Err("gen requires at least one operation");
                for (auto &v : it->second) c.cos.push_back(parse_si_i64(v));
                raw.erase(it);
                auto len = raw.find("len");
                if (len != raw.end()) {
                    if (len->second.size() != 1) throw Err("len requires exactly one value");
                    c.seconds = parse_si_f64(len->second[0]);
                    raw.erase(len);
                }
                if (!raw.empty()) throw Err("invalid flags: [\"" + raw.begin()->first + "\"]");
                c.sample_rate = parse_si_u64(next_arg(a, &i, "sample rate argument required"));
            } else if (cmd == "ui") {
                c.op = Op::Ui;
            } else if (cmd == "eui") {
                c.op = Op::Eui;
                if (i < a.size()) c.filename = a[i++];
            } else {
                throw Err("unrecognised command");
            }
        } catch (const Err &e) {
            throw Err("processing command: \"" + cmd + "\"\n\nCaused by:\n    " + e.what());
        }
        out.push_back(c);
    }
    return out;
}

const char *fmt_name(int f)
{
    switch (f) {
    case QD_FMT_CF32: return "cf32";
    case QD_FMT_CS8: return "cs8";
    case QD_FMT_CU8: return "cu8";
    default: return "cs16";
    }
}

void dump(const std::vector<Command> &cmds)
{
    for (auto &c : cmds) {
        switch (c.op) {
        case Op::From: printf("From filename=%s format=%s sample_rate=%" PRIu64 "\n", c.filename.c_str(), fmt_name(c.format), c.sample_rate); break;
        case Op::Shift: printf("Shift frequency=%" PRId64 "\n", c.frequency); break;
        case Op::LowPass: printf("LowPass size=%" PRIu64 " decimate=%" PRIu64 " frequency=%" PRId64 "\n", c.size, c.decimate, c.frequency); break;
        case Op::SparkFft:
            printf("SparkFft width=%" PRIu64 " stride=%" PRIu64, c.width, c.stride);
            if (c.min) printf(" min=%.9g max=%.9g", *c.min, *c.max);
            printf("\n");
            break;
        case Op::Bucket: printf("Bucket fft_width=%" PRIu64 " stride=%" PRIu64 " levels=%" PRIu64 "\n", c.width, c.stride, c.levels); break;
        case Op::Write: printf("Write overwrite=%s prefix=%s\n", c.overwrite ? "true" : "false", c.prefix.c_str()); break;
        case Op::Gen:
            printf("Gen sample_rate=%" PRIu64 " seconds=%.17g cos=", c.sample_rate, c.seconds);
            for (size_t k = 0; k < c.cos.size(); k++) printf("%s%" PRId64, k ? "," : "", c.cos[k]);
            printf("\n");
            break;
        case Op::Ui: printf("Ui\n"); break;
        case Op::Eui: printf("Eui filename=%s\n", c.filename.c_str()); break;
        }
    }
}

// ---- execution: the fold of src/bin/quadrs.rs:48-56 over Operation::exec (src/lib.rs:83-175) ----
struct Graph {
    bool has_source = false;
    qd_source src{};
    std::string path;
    std::vector<int64_t> cos;
    std::vector<qd_stage> stages;
    qd_chain *chain = nullptr;
    bool dirty = true;
    std::vector<int> devices{0}; // --gpus N: the chain is sharded over devices 0..N-1 inside this one process

    Graph() = default;
    Graph(const Graph &) = delete;
    Graph &operator=(const Graph &) = delete;
    ~Graph()
    {
        if (chain) qd_chain_destroy(chain);
    }
    // a new `from` / `gen` starts a new graph (lib.rs:89-101 ignore the previous value)
    void reset()
    {
        if (chain) qd_chain_destroy(chain);
        chain = nullptr;
        has_source = false;
        src = qd_source{};
        path.clear();
        cos.clear();
        stages.clear();
        dirty = true;
    }
    // stage constructors validate eagerly, as Shift::new / LowPass::new do
    void rebuild()
    {
        if (chain) qd_chain_destroy(chain);
        chain = nullptr;
        src.path = path.c_str();
        src.gen_cos = cos.data();
        src.gen_n_cos = cos.size();
        const int rc = qd_chain_create_sharded(&src, stages.data(), stages.size(), devices.data(), devices.size(), &chain);
        if (rc != QD_OK) throw Err(qd_last_error());
        dirty = false;
    }
};

void check(int rc)
{
    if (rc != QD_OK) throw Err(qd_last_error());
}

void exec(Graph &g, const Command &c)
{
    switch (c.op) {
    case Op::From:
        g.reset();
        g.has_source = true;
        g.src.kind = QD_SRC_FILE;
        g.src.format = c.format;
        g.src.sample_rate = c.sample_rate;
        g.path = c.filename;
        g.rebuild();
        break;
    case Op::Gen:
        g.reset();
        g.has_source = true;
        g.src.kind = QD_SRC_GEN;
        g.src.sample_rate = c.sample_rate;
        g.src.gen_seconds = c.seconds;
        g.cos = c.cos;
        g.rebuild();
        break;
    case Op::Shift: {
        if (!g.has_source) throw Err("shift requires an input");
        qd_stage s{};
        s.kind = QD_STAGE_SHIFT;
        s.frequency = c.frequency;
        g.stages.push_back(s);
        g.rebuild();
        break;
    }
    case Op::LowPass: {
        if (!g.has_source) throw Err("lowpass requires an input");
        qd_stage s{};
        s.kind = QD_STAGE_LOWPASS;
        s.frequency = c.frequency;
        s.decimate = c.decimate;
        s.size = c.size;
        g.stages.push_back(s);
        g.rebuild();
        break;
    }
    case Op::SparkFft: { // fft::spark_fft, src/fft.rs:12-69
        if (!g.has_source) throw Err("sparkfft requires an input");
        uint64_t rate = 0, rows = 0;
        check(qd_chain_sample_rate(g.chain, &rate));
        printf("sparkfft sample_rate=%" PRIu64 "\n", rate); // fft.rs:19, before anything can fail
        fflush(stdout);
        int rc = qd_sparkfft_rows(g.chain, c.width, c.stride, &rows);
        if (rc != QD_OK) throw Err(qd_last_error());
        if (rows == 0) rows = 1; // len <= width: the reference still attempts the first read
        const uint64_t batch = std::max<uint64_t>(1, (uint64_t(64) << 20) / std::max<uint64_t>(1, c.width));
        std::vector<uint8_t> idx(static_cast<size_t>(std::min(rows, batch) * c.width));
        std::vector<char> line(3 * c.width + 16);
        for (uint64_t r0 = 0; r0 < rows; r0 += batch) {
            uint64_t got = 0;
            rc = qd_sparkfft(g.chain, c.width, c.stride, c.min.has_value(), c.min.value_or(0), c.max.value_or(0), r0,
                             std::min(batch, rows - r0), idx.data(), nullptr, QD_SPACE_HOST, &got);
            // fft.rs:59 panics while it formats the first row that holds an out-of-range bin: the rows before it
            // are on stdout, that row and everything after it never appear
            uint64_t lim = got;
            if (rc == QD_E_GLYPH_RANGE)
                for (uint64_t r = 0; r < got && lim == got; r++)
                    for (uint64_t b = 0; b < c.width; b++)
                        if (idx[r * c.width + b] == 9) {
                            lim = r;
                            break;
                        }
            for (uint64_t r = 0; r < lim; r++) {
                const size_t n = qd_format_row(idx.data() + r * c.width, c.width, line.data(), line.size());
                fwrite(line.data(), 1, n, stdout);
                fputc('\n', stdout);
            }
            if (rc == QD_E_GLYPH_RANGE) {
                fflush(stdout);
                fprintf(stderr, "thread 'main' panicked at src/fft.rs:59: %s\n", qd_last_error());
                exit(101);
            }
            if (rc != QD_OK) throw Err(qd_last_error());
            if (got == 0) break;
        }
        break;
    }
    case Op::Bucket: { // lib.rs:139-160 + fft::freq_levels
        if (!g.has_source) throw Err("bucket -by freq requires an input");
        uint64_t total = 0;
        check(qd_freq_levels(g.chain, c.width, c.stride, c.levels, 0, 0, nullptr, QD_SPACE_HOST, &total));
        std::vector<uint8_t> vals(static_cast<size_t>(total ? total : 1));
        check(qd_freq_levels(g.chain, c.width, c.stride, c.levels, 0, total, vals.data(), QD_SPACE_HOST, &total));
        std::string s;
        for (uint64_t k = 0; k < total; k++) s += static_cast<char>('0' + vals[k]);
        printf("%s\n", s.c_str());
        break;
    }
    case Op::Write: { // do_write, src/lib.rs:178-213
        if (!g.has_source) throw Err("write requires an input");
        char name[4096];
        const int rc = qd_write_file(g.chain, c.prefix.c_str(), c.overwrite, name, sizeof name);
        if (rc == QD_E_WRITE_SHORT) {
            // the reference panics here too (lib.rs:203), after every readable sample is on disk
            fprintf(stderr, "thread 'main' panicked at src/lib.rs:203: %s\n", qd_last_error());
            exit(101);
        }
        check(rc);
        break;
    }
    case Op::Ui:
    case Op::Eui: throw Err("the ui / eui viewers are not part of the GPU build");
    }
}

} // namespace

int main(int argc, char **argv)
{
    std::vector<std::string> args(argv + 1, argv + argc);
    bool parse_only = false;
    int n_gpus = 1;
    // options of this front end only (the reference has none); they precede the commands
    while (!args.empty() && args[0].rfind("--", 0) == 0) {
        if (args[0] == "--parse-only") {
            parse_only = true;
            args.erase(args.begin());
        } else if (args[0] == "--gpus" && args.size() >= 2) {
            n_gpus = atoi(args[1].c_str());
            args.erase(args.begin(), args.begin() + 2);
            if (n_gpus < 1 || n_gpus > 64) {
                fprintf(stderr, "Error: --gpus takes a device count between 1 and 64\n");
                return 1;
            }
        } else {
            break;
        }
    }
    std::vector<Command> cmds;
    try {
        cmds = parse(args);
    } catch (const Err &e) {
        usage(argv[0]);
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    if (cmds.empty()) {
        usage(argv[0]);
        fprintf(stderr, "Error: no commands provided\n");
        return 1;
    }
    if (parse_only) {
        dump(cmds);
        return 0;
    }
    try {
        Graph g;
        g.devices.clear();
        for (int d = 0; d < n_gpus; d++) g.devices.push_back(d);
        for (auto &c : cmds) exec(g, c);
    } catch (const Err &e) {
        fflush(stdout);
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
