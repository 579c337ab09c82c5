// Deterministic synthetic Kubernetes objects for the benchmark / parity workloads.
// Built as its own small library (libgk_synth.so): workload generation is not part of the engine.
//
// Every object is a pure function of (seed, index) through splitmix64, so any index range can be generated
// independently (sharding across ranks, bounded CPU-baseline samples) and always yields the same bytes.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
  bool p(double q) { return uni() < q; }
};

constexpr int kNamespaces = 1000;
const char* kSpecialNs[] = {"kube-system", "gatekeeper-system", "production"};
const char* kRegistries[] = {"openpolicyagent/", "gcr.io/proj-%02u/", "docker.io/library/", "quay.io/", "registry.k8s.io/", "evil.example.com/"};
const char* kCpu[] = {"\"100m\"", "\"250m\"", "\"500m\"", "\"1\"", "\"2\"", "4"};
const char* kMem[] = {"\"64Mi\"", "\"128Mi\"", "\"512Mi\"", "\"1Gi\"", "\"2G\"", "\"4Gi\""};
const char* kHostPaths[] = {"/tmp", "/var/log", "/foo/bar", "/foo", "/etc", "/fool"};

struct ZipfNs {
  std::vector<double> cdf;
  ZipfNs() {
    double sum = 0;
    cdf.resize(kNamespaces);
    for (int i = 0; i < kNamespaces; ++i) {
      sum += 1.0 / std::pow((double)(i + 1), 1.1);
      cdf[i] = sum;
    }
    for (auto& c : cdf) c /= sum;
  }
  int pick(double u) const {
    int lo = 0, hi = kNamespaces - 1;
    while (lo < hi) {
      int mid = (lo + hi) / 2;
      if (cdf[mid] < u) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  }
};
const ZipfNs& zipf() {
  static ZipfNs z;
  return z;
}

void ns_name(Rng& r, std::string& out) {
  // 3% of objects land in the three well-known namespaces, the rest Zipf(1.1) over ns-0000..ns-0999
  double u = r.uni();
  if (u < 0.03) {
    out = kSpecialNs[r.below(3)];
    return;
  }
  char b[16];
  snprintf(b, sizeof b, "ns-%04d", zipf().pick(r.uni()));
  out = b;
}

void labels(Rng& r, std::string& o, bool want_team) {
  // 0-6 labels from 32 keys x 256 values; P(team) = 0.7
  o += "\"labels\":{";
  bool first = true;
  auto put = [&](const char* k, const char* v) {
    if (!first) o += ",";
    first = false;
    o += "\"";
    o += k;
    o += "\":\"";
    o += v;
    o += "\"";
  };
  char kb[24], vb[24];
  if (want_team && r.p(0.7)) {
    snprintf(vb, sizeof vb, "team-%u", r.below(64));
    put("team", vb);
  }
  uint32_t n = r.below(6);
  uint32_t used = 0;
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t k = r.below(32);
    if (used & (1u << k)) continue;
    used |= 1u << k;
    if (k == 0) snprintf(kb, sizeof kb, "app");
    else if (k == 1) snprintf(kb, sizeof kb, "env");
    else if (k == 2) snprintf(kb, sizeof kb, "owner");
    else if (k == 3) snprintf(kb, sizeof kb, "tier");
    else snprintf(kb, sizeof kb, "label-%02u", k);
    if (k == 1) snprintf(vb, sizeof vb, "%s", r.p(0.5) ? "prod" : (r.p(0.5) ? "staging" : "dev"));
    else snprintf(vb, sizeof vb, "v%u", r.below(256));
    put(kb, vb);
  }
  o += "}";
}

void container(Rng& r, std::string& o, uint32_t ci, bool init, uint32_t nvol) {
  char b[96];
  o += "{\"name\":\"";
  snprintf(b, sizeof b, "%s%u", init ? "init" : "c", ci);
  o += b;
  o += "\",\"image\":\"";
  uint32_t reg = r.below(100);
  // registry mix: 8% from a registry no allow-list contains
  uint32_t ri = reg < 30 ? 0 : reg < 55 ? 1 : reg < 72 ? 2 : reg < 84 ? 3 : reg < 92 ? 4 : 5;
  if (ri == 1) snprintf(b, sizeof b, kRegistries[1], r.below(20));
  else snprintf(b, sizeof b, "%s", kRegistries[ri]);
  o += b;
  snprintf(b, sizeof b, "repo-%03u", r.below(200));
  o += b;
  double t = r.uni();
  if (t < 0.15) o += ":latest";
  else if (t < 0.80) {
    snprintf(b, sizeof b, ":v%u.%u.%u", r.below(4), r.below(10), r.below(20));
    o += b;
  } else if (t < 0.97) {
    snprintf(b, sizeof b, ":sha-%08x", (uint32_t)r.next());
    o += b;
  }  // else: no tag at all
  o += "\"";
  double sc = r.uni();
  if (sc >= 0.60) {
    o += ",\"securityContext\":{\"privileged\":";
    o += sc >= 0.97 ? "true" : "false";
    if (r.p(0.3)) o += ",\"runAsNonRoot\":true";
    o += "}";
  }
  double res = r.uni();
  if (res < 0.80) {
    o += ",\"resources\":{\"limits\":{";
    bool cpu = r.p(0.93), mem = r.p(0.93);
    if (cpu) {
      o += "\"cpu\":";
      o += kCpu[r.below(6)];
    }
    if (mem) {
      if (cpu) o += ",";
      o += "\"memory\":";
      o += kMem[r.below(6)];
    }
    o += "}}";
  } else if (res < 0.88) {
    o += ",\"resources\":{}";
  }
  if (!init) {
    if (r.p(0.7)) o += r.p(0.8) ? ",\"readinessProbe\":{\"httpGet\":{\"path\":\"/healthz\",\"port\":8080}}" : ",\"readinessProbe\":{\"initialDelaySeconds\":5}";
    if (r.p(0.7)) o += r.p(0.5) ? ",\"livenessProbe\":{\"tcpSocket\":{\"port\":8080}}" : ",\"livenessProbe\":{\"exec\":{\"command\":[\"true\"]}}";
  }
  if (r.p(0.5)) {
    o += ",\"ports\":[{\"containerPort\":";
    snprintf(b, sizeof b, "%u", 1024 + r.below(8000));
    o += b;
    if (r.p(0.10)) {
      snprintf(b, sizeof b, ",\"hostPort\":%u", 1 + r.below(65535));
      o += b;
    }
    o += "}]";
  }
  if (nvol && r.p(0.6)) {
    o += ",\"volumeMounts\":[";
    uint32_t k = 1 + r.below(nvol);
    for (uint32_t i = 0; i < k; ++i) {
      if (i) o += ",";
      snprintf(b, sizeof b, "{\"name\":\"vol-%u\",\"mountPath\":\"/mnt/%u\"", r.below(nvol), i);
      o += b;
      double ro = r.uni();
      if (ro < 0.4) o += ",\"readOnly\":true";
      else if (ro < 0.5) o += ",\"readOnly\":false";
      o += "}";
    }
    o += "]";
  }
  o += "}";
}

void pod_spec(Rng& r, std::string& o) {
  char b[96];
  uint32_t nvol = 0;
  {
    double v = r.uni();
    nvol = v < 0.35 ? 0 : v < 0.70 ? 1 : v < 0.90 ? 2 : 3;
  }
  o += "\"spec\":{";
  double c = r.uni();
  uint32_t nc = c < 0.60 ? 1 : c < 0.85 ? 2 : c < 0.95 ? 3 : 4;
  o += "\"containers\":[";
  for (uint32_t i = 0; i < nc; ++i) {
    if (i) o += ",";
    container(r, o, i, false, nvol);
  }
  o += "]";
  if (r.p(0.2)) {
    o += ",\"initContainers\":[";
    container(r, o, 0, true, nvol);
    o += "]";
  }
  if (r.p(0.01)) o += ",\"hostPID\":true";
  if (r.p(0.01)) o += ",\"hostIPC\":true";
  if (r.p(0.01)) o += ",\"hostNetwork\":true";
  else if (r.p(0.02)) o += ",\"hostNetwork\":false";
  if (nvol) {
    o += ",\"volumes\":[";
    for (uint32_t i = 0; i < nvol; ++i) {
      if (i) o += ",";
      snprintf(b, sizeof b, "{\"name\":\"vol-%u\",", i);
      o += b;
      double t = r.uni();
      if (t < 0.05) {
        o += "\"hostPath\":{\"path\":\"";
        o += kHostPaths[r.below(6)];
        o += "\"}";
      } else if (t < 0.40) o += "\"emptyDir\":{}";
      else if (t < 0.62) {
        snprintf(b, sizeof b, "\"configMap\":{\"name\":\"cm-%u\"}", r.below(50));
        o += b;
      } else if (t < 0.80) {
        snprintf(b, sizeof b, "\"secret\":{\"secretName\":\"s-%u\"}", r.below(50));
        o += b;
      } else if (t < 0.88) o += "\"projected\":{\"sources\":[]}";
      else if (t < 0.97) {
        snprintf(b, sizeof b, "\"persistentVolumeClaim\":{\"claimName\":\"pvc-%u\"}", r.below(50));
        o += b;
      } else o += "\"nfs\":{\"server\":\"nfs.example.com\",\"path\":\"/exports\"}";
      o += "}";
    }
    o += "]";
  }
  o += "}";
}

void meta(Rng& r, std::string& o, const char* prefix, uint64_t idx, bool namespaced) {
  char b[64];
  o += "\"metadata\":{";
  if (r.p(0.03)) {
    snprintf(b, sizeof b, "\"generateName\":\"%s-gen-%u-\",", prefix, r.below(100));
    o += b;
  }
  snprintf(b, sizeof b, "\"name\":\"%s-%llu-%04x\"", prefix, (unsigned long long)idx, (unsigned)(r.next() & 0xffff));
  o += b;
  if (namespaced) {
    std::string ns;
    ns_name(r, ns);
    o += ",\"namespace\":\"" + ns + "\"";
  }
  o += ",";
  labels(r, o, true);
  o += "}";
}

void make_pod(uint64_t seed, uint64_t idx, std::string& o) {
  Rng r(seed ^ (idx * 0xD6E8FEB86659FD93ull + 0x6A7E6B33ull));
  o += "{\"apiVersion\":\"v1\",\"kind\":\"Pod\",";
  meta(r, o, "pod", idx, true);
  o += ",";
  pod_spec(r, o);
  o += "}";
}

// config 4: Pod .60 / Deployment .15 / Service .10 / ConfigMap .10 / Ingress .0499 / Namespace .0001
void make_mixed(uint64_t seed, uint64_t idx, std::string& o) {
  Rng r(seed ^ (idx * 0xD6E8FEB86659FD93ull + 0x6A7E6B33ull));
  double k = r.uni();
  char b[96];
  if (k < 0.60) {
    o += "{\"apiVersion\":\"v1\",\"kind\":\"Pod\",";
    meta(r, o, "pod", idx, true);
    o += ",";
    pod_spec(r, o);
    o += "}";
  } else if (k < 0.75) {
    o += "{\"apiVersion\":\"apps/v1\",\"kind\":\"Deployment\",";
    meta(r, o, "deploy", idx, true);
    snprintf(b, sizeof b, ",\"spec\":{\"replicas\":%u,\"template\":{", 1 + r.below(5));
    o += b;
    o += "\"metadata\":{";
    labels(r, o, true);
    o += "},";
    pod_spec(r, o);
    o += "}}}";
  } else if (k < 0.85) {
    o += "{\"apiVersion\":\"v1\",\"kind\":\"Service\",";
    meta(r, o, "svc", idx, true);
    snprintf(b, sizeof b, ",\"spec\":{\"type\":\"%s\",\"ports\":[{\"port\":%u}]}}", r.p(0.8) ? "ClusterIP" : "LoadBalancer", 80 + r.below(9000));
    o += b;
  } else if (k < 0.95) {
    o += "{\"apiVersion\":\"v1\",\"kind\":\"ConfigMap\",";
    meta(r, o, "cm", idx, true);
    o += ",\"data\":{\"key\":\"value\"}}";
  } else if (k < 0.9999) {
    o += "{\"apiVersion\":\"networking.k8s.io/v1\",\"kind\":\"Ingress\",";
    meta(r, o, "ing", idx, true);
    snprintf(b, sizeof b, ",\"spec\":{\"rules\":[{\"host\":\"h%u.example.com\"}]}}", r.below(1000));
    o += b;
  } else {
    o += "{\"apiVersion\":\"v1\",\"kind\":\"Namespace\",";
    meta(r, o, "ns", idx, false);
    o += "}";
  }
}

void make_namespace(uint64_t seed, int i, std::string& o) {
  Rng r(seed ^ (0xABCDEF12345ull + (uint64_t)i * 0x9E3779B97F4A7C15ull));
  char b[32];
  if (i < kNamespaces) snprintf(b, sizeof b, "ns-%04d", i);
  else snprintf(b, sizeof b, "%s", kSpecialNs[i - kNamespaces]);
  o += "{\"apiVersion\":\"v1\",\"kind\":\"Namespace\",\"metadata\":{\"name\":\"";
  o += b;
  o += "\",\"labels\":{";
  uint32_t n = r.below(5);
  bool first = true;
  auto put = [&](const char* k, const char* v) {
    if (!first) o += ",";
    first = false;
    o += std::string("\"") + k + "\":\"" + v + "\"";
  };
  if (n > 0) put("env", r.p(0.4) ? "prod" : (r.p(0.5) ? "staging" : "dev"));
  if (n > 1) put("tier", r.p(0.5) ? "frontend" : "backend");
  if (n > 2) put("admission.gatekeeper.sh/ignore", "no");
  if (n > 3) {
    char v[16];
    snprintf(v, sizeof v, "cc-%u", r.below(40));
    put("cost-center", v);
  }
  if (i >= kNamespaces) put("kubernetes.io/metadata.name", b);
  o += "}}}";
}

}  // namespace

extern "C" {

// mode 0: Pods (config 2/5), mode 1: mixed GVK (config 4).  *buf holds the concatenated documents,
// (*offsets)[count+1] the document boundaries.  Free both with gk_synth_free.
int gk_synth_objects(uint64_t seed, uint64_t start, uint64_t count, uint32_t mode, int threads, char** buf, uint64_t** offsets) {
  if (!buf || !offsets) return -1;
  int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
  if ((uint64_t)T > std::max<uint64_t>(1, count / 4096)) T = (int)std::max<uint64_t>(1, count / 4096);
  std::vector<std::string> parts(T);
  std::vector<std::vector<uint64_t>> lens(T);
  auto work = [&](int t) {
    uint64_t lo = count * t / T, hi = count * (t + 1) / T;
    std::string& s = parts[t];
    s.reserve((hi - lo) * 900);
    lens[t].reserve(hi - lo);
    for (uint64_t i = lo; i < hi; ++i) {
      size_t before = s.size();
      if (mode == 1) make_mixed(seed, start + i, s);
      else make_pod(seed, start + i, s);
      lens[t].push_back(s.size() - before);
    }
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  uint64_t total = 0;
  for (auto& p : parts) total += p.size();
  char* out = (char*)malloc(total + 1);
  uint64_t* off = (uint64_t*)malloc((count + 1) * sizeof(uint64_t));
  if (!out || !off) return -2;
  uint64_t pos = 0, k = 0;
  for (int t = 0; t < T; ++t) {
    memcpy(out + pos, parts[t].data(), parts[t].size());
    uint64_t p = pos;
    for (uint64_t l : lens[t]) {
      off[k++] = p;
      p += l;
    }
    pos += parts[t].size();
  }
  off[count] = pos;
  out[pos] = 0;
  *buf = out;
  *offsets = off;
  return 0;
}

int gk_synth_namespaces(uint64_t seed, char** buf, uint64_t** offsets, uint64_t* count) {
  std::string s;
  std::vector<uint64_t> off;
  int n = kNamespaces + 3;
  for (int i = 0; i < n; ++i) {
    off.push_back(s.size());
    make_namespace(seed, i, s);
  }
  off.push_back(s.size());
  char* out = (char*)malloc(s.size() + 1);
  uint64_t* o = (uint64_t*)malloc(off.size() * sizeof(uint64_t));
  if (!out || !o) return -2;
  memcpy(out, s.data(), s.size() + 1);
  memcpy(o, off.data(), off.size() * sizeof(uint64_t));
  *buf = out;
  *offsets = o;
  *count = (uint64_t)n;
  return 0;
}

void gk_synth_free(void* p) { free(p); }

}  // extern "C"
