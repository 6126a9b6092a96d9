//! `verlet`, `euler` and `rk4` integrator elements backed by libphysim_b200.so.
//!
//! physim loads integrators through an `extern "Rust"` constructor returning
//! `Box<dyn IntegratorElement>` (physim-core/src/plugin/mod.rs:92-103), so a C library cannot be an
//! integrator plugin by itself; this crate is the thin FFI layer (bindings as cbindgen would emit
//! them from include/physim_b200.h).  Behaviour mirrors integrators/src/verlet.rs:86-107.
//!
//! NOT compiled in the physim_b200 repository (no rustc in its build image).
use std::{collections::HashMap, ffi::c_void};

use physim_attribute::integrator_element;
use physim_core::{
    Acceleration, Entity,
    messages::MessageClient,
    plugin::{Element, ElementCreator, integrator::IntegratorElement},
    register_plugin,
};
use serde_json::Value;

register_plugin!("verlet", "euler", "rk4");

type AccFn = unsafe extern "C" fn(ctx: *mut c_void, state: *const Entity, n: usize, acc: *mut Acceleration);

unsafe extern "C" {
    // include/physim_b200.h: Pb200Integrator { PB200_VERLET = 0, PB200_EULER = 1, PB200_RK4 = 2 }
    fn pb200_integrator_create(kind: i32) -> *mut c_void;
    fn pb200_integrator_destroy(v: *mut c_void);
    fn pb200_integrator_step(
        v: *mut c_void,
        entities: *const Entity,
        new_state: *mut Entity,
        n: usize,
        acc_fn: AccFn,
        ctx: *mut c_void,
        dt: f64,
    ) -> i32;
    // the fused step (include/physim_b200.h "Fused step"): accelerations never leave the device
    fn pb200_verlet_create() -> *mut c_void;
    fn pb200_verlet_destroy(v: *mut c_void);
    fn pb200_verlet_set_resident(v: *mut c_void, on: i32) -> i32;
    fn pb200_verlet_step_fused(v: *mut c_void, transform: *mut c_void, entities: *const Entity, new_state: *mut Entity, n: usize, dt: f64) -> i32;
    // Pb200Kind { PB200_ASTRO = 0, PB200_ASTRO2 = 1, PB200_SIMPLE_ASTRO = 2 }; NaN = the element's default (1.0)
    fn pb200_transform_create(kind: i32, theta: f64, e: f64) -> *mut c_void;
    fn pb200_transform_destroy(t: *mut c_void);
}

#[integrator_element(
    name = "verlet",
    blurb = "Evaluate evolution with time using Verlet integration (B200)"
)]
struct Verlet {
    handle: *mut c_void,
    /// `verlet gravity=astro2 theta=1.5 e=0.5 [resident=true]`: the gravity element lives inside the integrator
    /// (pb200_verlet_step_fused) and is left out of the pipeline; null = the stock composition through acc_fn.
    fused: *mut c_void,
    gravity: *mut c_void,
}

unsafe impl Send for Verlet {}
unsafe impl Sync for Verlet {}

/// Turns the `&dyn Fn(&[Entity], &mut [Acceleration])` closure (pipeline.rs:137-141) into ctx + fn ptr.
unsafe extern "C" fn trampoline(ctx: *mut c_void, state: *const Entity, n: usize, acc: *mut Acceleration) {
    let f = unsafe { &*(ctx as *const &dyn Fn(&[Entity], &mut [Acceleration])) };
    let (s, a) = unsafe { (std::slice::from_raw_parts(state, n), std::slice::from_raw_parts_mut(acc, n)) };
    f(s, a)
}

impl IntegratorElement for Verlet {
    fn integrate(
        &self,
        entities: &[Entity],
        new_state: &mut [Entity],
        acc_fn: &dyn Fn(&[Entity], &mut [Acceleration]),
        dt: f64,
    ) {
        let ctx = &acc_fn as *const &dyn Fn(&[Entity], &mut [Acceleration]) as *mut c_void;
        let rc = unsafe {
            if !self.fused.is_null() {
                pb200_verlet_step_fused(self.fused, self.gravity, entities.as_ptr(), new_state.as_mut_ptr(), entities.len(), dt)
            } else {
                pb200_integrator_step(self.handle, entities.as_ptr(), new_state.as_mut_ptr(), entities.len(), trampoline, ctx, dt)
            }
        };
        if rc != 0 {
            eprintln!("physim_b200 verlet step failed");
            std::process::exit(1) // integrators/src/verlet.rs:98-100 exits on an unusable state too
        }
    }
}

impl Drop for Verlet {
    fn drop(&mut self) {
        unsafe {
            pb200_integrator_destroy(self.handle);
            if !self.fused.is_null() {
                pb200_verlet_destroy(self.fused);
                pb200_transform_destroy(self.gravity);
            }
        }
    }
}

impl MessageClient for Verlet {}

impl ElementCreator for Verlet {
    fn create_element(properties: HashMap<String, Value>) -> Box<Self> {
        let num = |k: &str| properties.get(k).and_then(|v| v.as_f64()).unwrap_or(f64::NAN);
        let kind = match properties.get("gravity").and_then(|v| v.as_str()) {
            Some("astro") => Some(0),
            Some("astro2") => Some(1),
            Some("simple_astro") => Some(2),
            _ => None,
        };
        let (mut fused, mut gravity) = (std::ptr::null_mut(), std::ptr::null_mut());
        if let Some(kind) = kind {
            unsafe {
                fused = pb200_verlet_create();
                gravity = pb200_transform_create(kind, num("theta"), num("e"));
                if properties.get("resident").and_then(|v| v.as_bool()).unwrap_or(false) {
                    pb200_verlet_set_resident(fused, 1);
                }
            }
        }
        Box::new(Self { handle: unsafe { pb200_integrator_create(0) }, fused, gravity })
    }
}

impl Element for Verlet {
    fn get_property_descriptions(&self) -> Result<HashMap<String, String>, Box<dyn std::error::Error>> {
        Ok(HashMap::from([
            (String::from("gravity"), String::from("astro | astro2 | simple_astro: evaluate this gravity element inside the integrator (accelerations stay on the GPU); leave that element out of the pipeline")),
            (String::from("theta"), String::from("Barnes-Hut parameter of the fused gravity element. Default=1.0")),
            (String::from("e"), String::from("Easing factor of the fused gravity element. Default=1.0")),
            (String::from("resident"), String::from("true: keep the state in GPU memory between steps; the input is checked on a sample against the previous output and uploaded only when it differs. Default=false")),
        ]))
    }
}

// `euler` (integrators/src/euler.rs) and `rk4` (integrators/src/rk4.rs): same forwarding, other kind.
macro_rules! forwarded_integrator {
    ($ty:ident, $name:literal, $blurb:literal, $kind:literal) => {
        #[integrator_element(name = $name, blurb = $blurb)]
        struct $ty {
            handle: *mut c_void,
        }
        unsafe impl Send for $ty {}
        unsafe impl Sync for $ty {}
        impl IntegratorElement for $ty {
            fn integrate(
                &self,
                entities: &[Entity],
                new_state: &mut [Entity],
                acc_fn: &dyn Fn(&[Entity], &mut [Acceleration]),
                dt: f64,
            ) {
                let ctx = &acc_fn as *const &dyn Fn(&[Entity], &mut [Acceleration]) as *mut c_void;
                let rc = unsafe {
                    pb200_integrator_step(self.handle, entities.as_ptr(), new_state.as_mut_ptr(), entities.len(), trampoline, ctx, dt)
                };
                if rc != 0 {
                    eprintln!("physim_b200 {} step failed", $name);
                    std::process::exit(1)
                }
            }
        }
        impl Drop for $ty {
            fn drop(&mut self) {
                unsafe { pb200_integrator_destroy(self.handle) }
            }
        }
        impl MessageClient for $ty {}
        impl ElementCreator for $ty {
            fn create_element(_: HashMap<String, Value>) -> Box<Self> {
                Box::new(Self { handle: unsafe { pb200_integrator_create($kind) } })
            }
        }
        impl Element for $ty {
            fn get_property_descriptions(&self) -> Result<HashMap<String, String>, Box<dyn std::error::Error>> {
                Ok(HashMap::from([]))
            }
        }
    };
}
forwarded_integrator!(Euler, "euler", "Evaluate evolution with time using Euler integration (B200)", 1);
forwarded_integrator!(Rk4, "rk4", "Evaluate evolution with time using Rk4 integration (B200)", 2);
