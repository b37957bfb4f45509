// tests/native/test_setters_tsan.cc -- TEST INFRASTRUCTURE.  The real host engine (engine.cc) on the emulated device
// (device_emu.cc), built with -fsanitize=thread: four threads hammer every queued setter of the C ABI (positions, volumes,
// occlusions, head, master volume, mix parameters -- also for ids that are being destroyed) while the mixer thread runs
// sbm_mix_submit / sbm_mix_collect with two buffers in flight and churns sources.  Prints the worst and median setter
// latency; ThreadSanitizer reports (if any) go to stderr and make the exit code non-zero.
// Reference behaviour matched: Resonance's setters post to a lock-free queue and never block
// (RA/utils/lockless_task_queue.cc:43-51, RA/graph/resonance_audio_api_impl.cc:305-317).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/sbm_audio.h"

int main() {
	sbm_config cfg{};
	cfg.frames_per_buffer = 256; cfg.sample_rate_hz = 48000; cfg.device = 0; cfg.max_sources = 512; cfg.max_mix_sources = 512;
	sbm_engine* e = nullptr;
	if (sbm_create(&cfg, &e) != SBM_OK) return 2;
	std::vector<float> pcm(4096);
	for (size_t i = 0; i < pcm.size(); ++i) pcm[i] = 0.25f * (float)((i * 37) % 101) / 101.0f - 0.1f;
	int buf = -1;
	if (sbm_buffer_create(e, pcm.data(), pcm.size(), &buf) != SBM_OK) return 2;
	const int N = 256;
	int first = 0;
	sbm_source_create(e, SBM_SOURCE_STEREO_PANNING, N, &first);
	std::vector<int> ids(N), bufs(N, buf);
	std::vector<double> pos(N, 0.0), delta(N, 1.01);
	std::vector<float> mixf(N, 0.8f);
	for (int i = 0; i < N; ++i) ids[i] = first + i;
	sbm_source_attach_mix1(e, N, ids.data(), bufs.data(), pos.data(), delta.data(), mixf.data());

	std::atomic<bool> stop{false};
	std::atomic<int> max_id{first + N};
	std::vector<std::vector<double>> lat(4);
	std::vector<std::thread> th;
	for (int t = 0; t < 4; ++t)
		th.emplace_back([&, t]() {
			unsigned s = 12345u + 977u * (unsigned)t;
			auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
			std::vector<int> my(64); std::vector<float> xyz(192), v(64); std::vector<int> z(64, 0); std::vector<double> d(64);
			while (!stop.load(std::memory_order_relaxed)) {
				const int hi = max_id.load(std::memory_order_relaxed);
				for (int k = 0; k < 64; ++k) {
					my[k] = (int)(rnd() % (unsigned)hi);
					xyz[3 * k] = (float)(rnd() % 80) - 40.0f; xyz[3 * k + 1] = (float)(rnd() % 20) - 10.0f; xyz[3 * k + 2] = (float)(rnd() % 80) - 40.0f;
					v[k] = 0.25f + 0.75f * (float)(rnd() % 100) / 100.0f; d[k] = 0.95 + 0.1 * (double)(rnd() % 100) / 100.0;
				}
				const float hp[3] = {(float)(rnd() % 10), 0.0f, 0.0f}, hq[4] = {0.0f, 0.0f, 0.0f, 1.0f};
				const auto t0 = std::chrono::steady_clock::now();
				switch (rnd() % 6) {
					case 0: sbm_set_source_positions(e, 64, my.data(), xyz.data()); break;
					case 1: sbm_set_source_volumes(e, 64, my.data(), v.data()); break;
					case 2: sbm_set_source_occlusions(e, 64, my.data(), v.data()); break;
					case 3: sbm_set_head(e, hp, hq); break;
					case 4: sbm_set_master_volume(e, 0.9f); break;
					default: sbm_source_set_mix_params(e, 64, my.data(), z.data(), d.data(), v.data()); break;
				}
				lat[(size_t)t].push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
				std::this_thread::sleep_for(std::chrono::microseconds(50));
			}
		});

	std::vector<float> out(512);
	int valid = 0, rc = 0;
	unsigned s = 99;
	const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(1500);
	int buffers = 0;
	while (std::chrono::steady_clock::now() < t_end && !rc) {
		rc |= sbm_mix_submit(e);
		if (buffers >= 1) rc |= sbm_mix_collect(e, out.data(), &valid);
		// churn: destroy 8 random sources, create 8 new ones (ids keep increasing, slots and MixZ entries are recycled)
		int kill[8];
		for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; kill[k] = (int)((s >> 8) % (unsigned)max_id.load()); }
		sbm_source_destroy(e, 8, kill);     // unknown / already dead ids: SBM_ERR_UNKNOWN_ID, ignored here
		int nf = 0;
		if (sbm_source_create(e, SBM_SOURCE_STEREO_PANNING, 8, &nf) == SBM_OK) {
			int nid[8]; for (int k = 0; k < 8; ++k) nid[k] = nf + k;
			rc |= sbm_source_attach_mix1(e, 8, nid, bufs.data(), pos.data(), delta.data(), mixf.data());
			max_id.store(nf + 8);
		}
		++buffers;
	}
	rc |= sbm_mix_collect(e, out.data(), &valid);
	stop = true;
	for (auto& x : th) x.join();
	std::vector<double> all;
	for (auto& l : lat) all.insert(all.end(), l.begin(), l.end());
	std::sort(all.begin(), all.end());
	sbm_stats st;
	sbm_get_stats(e, &st);
	printf("buffers %d setter_calls %zu median_us %.2f p999_us %.2f worst_us %.2f ignored_items %u rc %d\n", buffers, all.size(), all[all.size() / 2],
	       all[(size_t)((double)all.size() * 0.999)], all.back(), st.ignored_setter_items, rc);
	sbm_destroy(e);
	return rc ? 1 : 0;
}
