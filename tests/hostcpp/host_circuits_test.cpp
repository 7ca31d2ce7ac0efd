// The circuit classes of host/halo2_b200.hpp (DelayEncryptCircuit, RSACircuit, PoseidonEncCircuit) - CPU only, no context:
// reads the inputs test_host_cpp.py wrote, synthesizes, and writes outputs / advice / counts back for comparison with the Python
// mirror of the same front-end.  usage: host_circuits_test <dir>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "halo2_b200.hpp"

template <typename T>
static std::vector<T> read_all(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> v(raw.size() / sizeof(T));
    std::memcpy(v.data(), raw.data(), v.size() * sizeof(T));
    return v;
}
template <typename T>
static void write_all(const std::string& path, const std::vector<T>& v) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    try {
        const auto n = read_all<uint8_t>(dir + "/n.bin"), e = read_all<uint8_t>(dir + "/e.bin"), x = read_all<uint8_t>(dir + "/x.bin");
        const auto key = read_all<de_fr>(dir + "/key.bin");
        const std::vector<de_fr> zero_message(2);
        // DelayEncryptCircuit: keygen pass, then the witness pass on three threads into a dirty buffer
        halo2_b200::DelayEncryptCircuit delay(n, e, x, zero_message);
        halo2_b200::Assignment a = delay.synthesize(16);
        write_all(dir + "/delay_outputs.bin", a.outputs());
        std::vector<de_fr> adv(size_t(5) << 16);
        std::memset(adv.data(), 0x5A, adv.size() * sizeof(de_fr));
        const de_assignment_info_t info = delay.witness(16, adv.data(), 3);
        int fails = 0;
        for (uint32_t c = 0; c < a.n_advice(); c++) {
            const auto col = a.advice(c);
            if (std::memcmp(col.data(), adv.data() + (size_t(c) << 16), col.size() * sizeof(de_fr))) { std::cerr << "advice column " << c << " differs\n"; fails++; }
        }
        if (info.used_rows != a.used_rows() || a.n_fixed() != 15 || a.n_advice() != 5) { std::cerr << "shape mismatch\n"; fails++; }
        write_all(dir + "/delay_advice0.bin", a.advice(0));
        write_all(dir + "/delay_copies.bin", a.copies());
        std::ofstream(dir + "/delay_rows.txt") << a.used_rows() << "\n";
        // RSACircuit
        halo2_b200::RSACircuit rsa(n, e, x);
        write_all(dir + "/rsa_outputs.bin", rsa.synthesize(17).outputs());
        // PoseidonEncCircuit
        halo2_b200::PoseidonEncCircuit pose(key[0], key[1], zero_message);
        halo2_b200::Assignment p = pose.synthesize(11);
        write_all(dir + "/pose_outputs.bin", p.outputs());
        write_all(dir + "/pose_fixed3.bin", p.fixed(3));
        // failure behaviour: too few rows is an exception carrying the front-end's message
        bool threw = false;
        try { delay.synthesize(12); } catch (const std::runtime_error& err) { threw = std::string(err.what()).find("not enough rows") != std::string::npos; }
        if (!threw) { std::cerr << "k = 12 did not fail\n"; fails++; }
        std::cout << (fails ? "FAIL" : "OK") << std::endl;
        return fails ? 1 : 0;
    } catch (const std::exception& err) {
        std::cerr << "exception: " << err.what() << std::endl;
        return 3;
    }
}
