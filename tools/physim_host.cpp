// physim_host.cpp — a minimal stand-in for the physim executable, in C++, used to exercise
// libphysim_b200.so exactly the way physim's Rust host does, without Rust:
//
//   discovery      physim-core/src/plugin/discover.rs:248-387  (abi info, register_plugin,
//                  {el}_register with the host allocator, throw-away instance with "{}")
//   loading        physim-core/src/plugin/transform.rs:58-83   ({el}_get_api, init(json, len))
//   bus            physim-core/src/plugin/mod.rs:234-244, messages.rs:201-226 (set_callback_target,
//                  post_bus_callback exported by the host, post_configuration_messages)
//   simulation     physim-core/src/pipeline.rs:137-182         (acc_fn over the transforms,
//                  integrator.integrate(state, new_state, acc_fn, dt), state = new_state.clone())
//
// The integrator is `verlet` through pb200_verlet_step (what rust_shim/ forwards to).
//
//   renderer       utilities/src/csvsink.rs:44-80 through pb200_csvsink_* (optional): receives the
//                  initial state, then every new state (pipeline.rs:129-131,179)
//
//   physim_host <lib.so> <element> <json-properties> <state.bin> <dt> <iterations> <out.bin> [<out.csv> <print_n>]
//
// Build: g++ -O2 -std=c++17 -rdynamic tools/physim_host.cpp -o physim_host -ldl
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/physim_b200.h"

static int g_bus_messages = 0;
static std::string g_last_topic, g_last_message;

extern "C" {
// the host symbols a plugin may resolve (messages.rs:201-226, plugin/mod.rs host_alloc_string)
void post_bus_callback(void* target, CMessage message) {
  if (!target || !message.topic || !message.message) {
    std::fprintf(stderr, "Error, message contents are null\n");
    return;
  }
  ++g_bus_messages;
  g_last_topic = message.topic;      // origin == C: the host copies, it does not free
  g_last_message = message.message;
}
char* host_alloc_string(const char* s) {
  char* p = static_cast<char*>(std::malloc(std::strlen(s) + 1));
  std::strcpy(p, s);
  return p;
}
void host_free_string(char* s) { std::free(s); }
}

template <class F>
static F sym(void* lib, const std::string& name, bool required = true) {
  void* p = dlsym(lib, name.c_str());
  if (!p && required) {
    std::fprintf(stderr, "missing symbol %s\n", name.c_str());
    std::exit(2);
  }
  return reinterpret_cast<F>(p);
}

struct Loaded {
  const TransformElementAPI* api;
  void* obj;
};
struct AccCtx {
  std::vector<Loaded>* transforms;
};
// pipeline.rs:137-141: every transform, in order
static void acc_fn(void* ctx, const Entity* state, size_t n, Acceleration* acc) {
  for (const Loaded& t : *static_cast<AccCtx*>(ctx)->transforms) t.api->transform(t.obj, state, n, acc, n);
}

int main(int argc, char** argv) {
  if (argc != 8 && argc != 10) {
    std::fprintf(stderr, "usage: %s lib element json state.bin dt iterations out.bin [out.csv print_n]\n", argv[0]);
    return 2;
  }
  const std::string element = argv[2], props = argv[3];
  const double dt = std::atof(argv[5]);
  const long iterations = std::atol(argv[6]);
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    std::fprintf(stderr, "dlopen: %s\n", dlerror());
    return 2;
  }
  // ---- discovery -------------------------------------------------------------------------
  const char* abi = sym<const char* (*)()>(lib, "get_plugin_abi_info")();
  if (std::strcmp(abi, "C") != 0) std::fprintf(stderr, "warning: ABI %s\n", abi);
  std::string list = sym<const char* (*)()>(lib, "register_plugin")();
  std::printf("plugin abi=%s elements=%s\n", abi, list.c_str());
  size_t pos = 0;
  bool found = false;
  while (pos <= list.size()) {
    const size_t comma = list.find(',', pos);
    const std::string el = list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
    pos = comma == std::string::npos ? list.size() + 1 : comma + 1;
    if (el.empty()) continue;
    ElementMetaFFI meta = sym<ElementMetaFFI (*)(RustStringAllocFn)>(lib, el + "_register")(host_alloc_string);
    // discover.rs:376-387: a throw-away instance just to read the property docs
    const TransformElementAPI* api = sym<const TransformElementAPI* (*)()>(lib, el + "_get_api")();
    void* tmp = api->init(reinterpret_cast<const uint8_t*>("{}"), 2);
    char* docs = tmp ? api->get_property_descriptions(tmp, host_alloc_string) : nullptr;
    std::printf("  element %-12s kind=%d blurb=\"%s\" properties=%s\n", meta.name, int(meta.kind), meta.blurb,
                docs ? docs : "?");
    if (docs) host_free_string(docs);
    if (tmp) api->destroy(tmp);
    for (char* s : {meta.name, meta.plugin, meta.version, meta.license, meta.author, meta.blurb, meta.repo})
      host_free_string(s);
    found |= (el == element);
  }
  if (!found) {
    std::fprintf(stderr, "element %s not registered\n", element.c_str());
    return 2;
  }
  // ---- pipeline build (PipelineBuilder::add) ------------------------------------------------
  static int bus_token = 0;
  sym<void (*)(void*)>(lib, "set_callback_target")(&bus_token);
  std::vector<Loaded> transforms;
  const TransformElementAPI* api = sym<const TransformElementAPI* (*)()>(lib, element + "_get_api")();
  void* obj = api->init(reinterpret_cast<const uint8_t*>(props.data()), props.size());  // not NUL-terminated use
  if (!obj) {
    std::fprintf(stderr, "Failed to load transform element\n");
    return 3;
  }
  transforms.push_back({api, obj});
  api->post_configuration_messages(obj);
  std::printf("bus messages=%d last=%s/%s\n", g_bus_messages, g_last_topic.c_str(), g_last_message.c_str());
  auto verlet_create = sym<void* (*)()>(lib, "pb200_verlet_create");
  auto verlet_step = sym<int (*)(void*, const Entity*, Entity*, size_t, Pb200AccFn, void*, double)>(lib, "pb200_verlet_step");
  auto verlet_destroy = sym<void (*)(void*)>(lib, "pb200_verlet_destroy");
  // ---- state ---------------------------------------------------------------------------------
  FILE* f = std::fopen(argv[4], "rb");
  if (!f) return 2;
  std::fseek(f, 0, SEEK_END);
  const size_t n = size_t(std::ftell(f)) / sizeof(Entity);
  std::fseek(f, 0, SEEK_SET);
  std::vector<Entity> state(n), new_state(n);
  if (std::fread(state.data(), sizeof(Entity), n, f) != n) return 2;
  std::fclose(f);
  new_state = state;
  // ---- simulation thread (pipeline.rs:143-182) ---------------------------------------------
  void* verlet = verlet_create();
  AccCtx ctx{&transforms};
  void* sink = nullptr;
  auto sink_push = sym<int (*)(void*, const Entity*, size_t)>(lib, "pb200_csvsink_push");
  if (argc == 10) {
    sink = sym<void* (*)(const char*, size_t)>(lib, "pb200_csvsink_create")(argv[8], size_t(std::atol(argv[9])));
    if (!sink) return 5;
    if (sink_push(sink, state.data(), n) != 0) return 5;  // simulation_sender.send(state.clone())
  }
  for (long it = 0; it < iterations; ++it) {
    if (verlet_step(verlet, state.data(), new_state.data(), n, acc_fn, &ctx, dt) != 0) return 4;
    state = new_state;  // state = new_state.clone()
    if (sink && sink_push(sink, new_state.data(), n) != 0) return 5;  // simulation_sender.send(new_state.clone())
  }
  if (sink) sym<void (*)(void*)>(lib, "pb200_csvsink_destroy")(sink);
  verlet_destroy(verlet);
  for (const Loaded& t : transforms) t.api->destroy(t.obj);
  f = std::fopen(argv[7], "wb");
  std::fwrite(state.data(), sizeof(Entity), n, f);
  std::fclose(f);
  std::printf("ran %ld iterations on %zu entities\n", iterations, n);
  dlclose(lib);
  return 0;
}
