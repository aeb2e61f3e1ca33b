"""The persistent host thread pool of snk_step_host_f64 (bullet_envs_b200/csrc/snake_hostpool.h), exercised without a GPU: a g++
harness widens float arrays through it many times, at sizes around the single-thread threshold, with ragged chunk counts."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include "snake_hostpool.h"
#include <stdio.h>
int main() {
    HostPool& p = HostPool::get();
    if (p.size() != 5) { printf("size %d\n", p.size()); return 2; }
    const size_t sizes[] = {1, 1000, 65535, 65536, 65537, 100003, 1u << 20, (1u << 20) + 7};
    for (int rep = 0; rep < 200; rep++)
        for (size_t n : sizes) {
            std::vector<float> in(n); std::vector<double> out(n, -1.0);
            for (size_t i = 0; i < n; i++) in[i] = (float)(i % 977) * 0.5f + rep;
            const float* a = in.data(); double* b = out.data();
            p.run(n, [=](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) b[i] = (double)a[i]; });
            for (size_t i = 0; i < n; i++) if (out[i] != (double)in[i]) { printf("mismatch n=%zu i=%zu\n", n, i); return 1; }
        }
    // two caller threads (one handle each in the library) share the pool: their jobs take turns
    std::vector<double> o1(300000, -1.0), o2(200000, -1.0);
    auto job = [&](std::vector<double>* o, double add) {
        for (int rep = 0; rep < 50; rep++) {
            double* b = o->data(); const size_t n = o->size();
            p.run(n, [=](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) b[i] = (double)i + add + rep; });
            for (size_t i = 0; i < n; i++) if (b[i] != (double)i + add + rep) { printf("concurrent mismatch\n"); exit(3); }
        }
    };
    std::thread t1(job, &o1, 0.5), t2(job, &o2, 0.25);
    t1.join(); t2.join();
    printf("ok\n");
    return 0;
}
'''


def test_host_pool_covers_every_element_exactly_once(tmp_path):
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "bullet_envs_b200", "csrc"), "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], env=dict(os.environ, SNK_HOST_THREADS="5"), stdout=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout
    # under torchrun the pool shrinks with the ranks on the node
    sized = "#include \"snake_hostpool.h\"\n#include <stdio.h>\nint main(){ printf(\"%d\\n\", HostPool::get().size()); return 0; }\n"
    (tmp_path / "size.cpp").write_text(sized)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "bullet_envs_b200", "csrc"), "-o", str(tmp_path / "size"), str(tmp_path / "size.cpp")])
    env = {k: v for k, v in os.environ.items() if k != "SNK_HOST_THREADS"}
    one = int(subprocess.run([str(tmp_path / "size")], env=dict(env, LOCAL_WORLD_SIZE="1"), stdout=subprocess.PIPE, text=True).stdout)
    eight = int(subprocess.run([str(tmp_path / "size")], env=dict(env, LOCAL_WORLD_SIZE="8"), stdout=subprocess.PIPE, text=True).stdout)
    assert 1 <= eight <= one <= 16 and eight <= max(1, (os.cpu_count() or 1) // 8)
