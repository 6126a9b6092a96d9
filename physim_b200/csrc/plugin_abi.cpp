// plugin_abi.cpp — the symbols physim's loader resolves in a plugin library, for the three gravity
// transforms.  Template: c_plugin/plugin.c (the reference's own C plugin); contract:
// physim-core/src/plugin/{discover,transform,meta,mod}.rs, physim-attribute/src/lib.rs:39-202.
//
// Element names, property names, defaults and blurbs are those of astro/src/transformers.rs so a
// pipeline such as `cube ... ! astro2 theta=1.5 e=0.5 ! verlet ! ...` needs no change.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/physim_b200.h"

namespace {

void* g_bus_target = nullptr;

const char kPlugin[] = "physim_b200";
const char kVersion[] = "0.1.0";
const char kLicense[] = "MIT";
const char kAuthor[] = "physim_b200 authors";
// this plugin has no public repository of its own; physim shows the string verbatim (meta.rs:95-135)
const char kRepo[] = "physim_b200 (drop-in CUDA plugin for physim; not the physim repository)";

// transformers.rs:91-104 / :182-196 / :258-263 (property docs shown by `physcan <element>`)
const char kBhProps[] =
    "{\"theta\":\"Barnes-Hut parameter. Increase for speed, decrease for accuracy. Default=1.0\","
    "\"e\":\"Easing factor. Modify G*Ma*Mb*(r-e)^-2. Default=1.0\"}";
const char kDirectProps[] = "{\"e\":\"Easing factor. Modify G*Ma*Mb*(r-e)^-2. Default=1.0\"}";

typedef void (*PostBusFn)(void*, CMessage);

ElementMetaFFI make_meta(RustStringAllocFn alloc, const char* name, const char* blurb) {
  ElementMetaFFI m;
  m.kind = Transform;
  m.name = alloc(name);
  m.plugin = alloc(kPlugin);
  m.version = alloc(kVersion);
  m.license = alloc(kLicense);
  m.author = alloc(kAuthor);
  m.blurb = alloc(blurb);
  m.repo = alloc(kRepo);
  return m;
}

// shared bodies of the per-element vtables -------------------------------------------------------

void transform_impl(const void* obj, const Entity* state, uintptr_t n, Acceleration* acc, uintptr_t n_acc) {
  if (!obj) return;
  // The C ABI has no error channel (transform.rs:85-104).  A failed GPU evaluation must not look
  // like a successful one with zero forces: report and abort, as the Rust trampolines do on panic
  // (physim-attribute/src/lib.rs:94-97).
  if (pb200_transform_apply(const_cast<void*>(obj), state, n, acc, n_acc) != 0) {
    std::fprintf(stderr, "[physim_b200] fatal: %s\n", pb200_last_error());
    std::abort();
  }
}

void destroy_impl(void* obj) { pb200_transform_destroy(obj); }

void recv_message_impl(void* obj, const CMessage* msg) {
  (void)obj;
  (void)msg;  // the reference elements ignore bus traffic (MessageClient default, messages.rs)
}

// transformers.rs:107-112,198-203,266-271: tell `energysink` that gravity is present
void post_configuration_impl(void* obj) {
  if (!obj || !g_bus_target) return;
  static PostBusFn post = reinterpret_cast<PostBusFn>(dlsym(RTLD_DEFAULT, "post_bus_callback"));
  if (!post) return;  // host symbol not visible (e.g. a test harness): nothing to tell
  CMessage m;
  m.priority = Low;
  m.topic = "energysink";
  m.message = "gravity";
  m.sender_id = reinterpret_cast<uintptr_t>(obj);
  m.origin = C;  // host copies the strings instead of freeing them (messages.rs:140-148)
  post(g_bus_target, m);
}

char* bh_props(void* obj, RustStringAllocFn alloc) { return (obj && alloc) ? alloc(kBhProps) : nullptr; }
char* direct_props(void* obj, RustStringAllocFn alloc) { return (obj && alloc) ? alloc(kDirectProps) : nullptr; }

void* astro_init(const uint8_t* json, uintptr_t len) { return pb200_transform_create_json(PB200_ASTRO, json, len); }
void* astro2_init(const uint8_t* json, uintptr_t len) { return pb200_transform_create_json(PB200_ASTRO2, json, len); }
void* simple_init(const uint8_t* json, uintptr_t len) {
  return pb200_transform_create_json(PB200_SIMPLE_ASTRO, json, len);
}

}  // namespace

extern "C" {

const char* get_plugin_abi_info(void) { return "C"; }

const char* register_plugin(void) { return "astro,astro2,simple_astro"; }

void set_callback_target(void* target) {
  if (target == nullptr) {
    std::fprintf(stderr, "Error: callback target is null\n");
    std::abort();
  }
  g_bus_target = target;
}

ElementMetaFFI astro_register(RustStringAllocFn alloc) {
  return make_meta(alloc, "astro",
                   "Compute approximate gravitational accelerations with the Barnes-Hut algorithm (quadtree)");
}
ElementMetaFFI astro2_register(RustStringAllocFn alloc) {
  return make_meta(alloc, "astro2",
                   "Compute approximate gravitational accelerations with the Barnes-Hut algorithm (octree)");
}
ElementMetaFFI simple_astro_register(RustStringAllocFn alloc) {
  return make_meta(alloc, "simple_astro", "Compute exact gravitational accelerations");
}

const TransformElementAPI* astro_get_api(void) {
  static const TransformElementAPI api = {astro_init,        transform_impl,        destroy_impl,
                                          bh_props,          recv_message_impl,     post_configuration_impl};
  return &api;
}
const TransformElementAPI* astro2_get_api(void) {
  static const TransformElementAPI api = {astro2_init,       transform_impl,        destroy_impl,
                                          bh_props,          recv_message_impl,     post_configuration_impl};
  return &api;
}
const TransformElementAPI* simple_astro_get_api(void) {
  static const TransformElementAPI api = {simple_init,       transform_impl,        destroy_impl,
                                          direct_props,      recv_message_impl,     post_configuration_impl};
  return &api;
}

}  // extern "C"
