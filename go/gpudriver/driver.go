// Package gpudriver is the Go side of the drop-in: a constraint-framework drivers.Driver whose Query and the
// additive ReviewBatch go through cgo into libgk_engine.so (include/gk_engine.h).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Go toolchain here (go: command not found) and the
// frameworks/constraint + OPA modules are not vendored.  The file is the binding a Gatekeeper maintainer would add;
// it follows the only in-tree Driver implementation line by line for locking, review type assertions and stats
// (pkg/drivers/k8scel/driver.go:60-263).  tests/ exercise the same C ABI through the ctypes mirror
// gatekeeper_b200/driver.py, method for method.
//
// Registration (replaces rego.New(args...) -- main.go:457-462, pkg/gator/opa.go:32-37, pkg/gator/test/test.go:48-53,
// pkg/gator/bench/bench.go:313-319):
//
//	d, err := gpudriver.New(gpudriver.Device(0))
//	client, err := constraintclient.NewClient(constraintclient.Targets(&target.K8sValidationTarget{}),
//	        constraintclient.Driver(k8scelDriver), constraintclient.Driver(d), constraintclient.EnforcementPoints(...))
package gpudriver

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gatekeeper_b200 -lgk_engine
#include <stdlib.h>
#include "gk_engine.h"
*/
import "C"

import (
	"context"
	"encoding/json"
	"errors"
	"fmt"
	"sync"
	"unsafe"

	"github.com/open-policy-agent/frameworks/constraint/pkg/client/drivers"
	"github.com/open-policy-agent/frameworks/constraint/pkg/client/reviews"
	"github.com/open-policy-agent/frameworks/constraint/pkg/core/templates"
	"github.com/open-policy-agent/frameworks/constraint/pkg/instrumentation"
	"github.com/open-policy-agent/frameworks/constraint/pkg/types"
	"github.com/open-policy-agent/opa/v1/storage"
	admissionv1 "k8s.io/api/admission/v1"
	"k8s.io/apimachinery/pkg/apis/meta/v1/unstructured"
)

// Name is "Rego": templates carrying targets[].rego / code[engine: Rego] route to this driver.  It replaces -- and
// cannot co-exist with -- rego.Driver (SURVEY.md 8(b)).
const Name = "Rego"

var _ drivers.Driver = &Driver{}

type Driver struct {
	mux sync.RWMutex // mutators exclusive, Query shared: pkg/drivers/k8scel/driver.go:61,130,167
	e   *C.gk_engine_t
	gatherStats bool
}

type Arg func(*Driver, *C.gk_cfg)

func Device(i int) Arg   { return func(_ *Driver, c *C.gk_cfg) { c.device = C.int32_t(i) } }
func Threads(n int) Arg  { return func(_ *Driver, c *C.gk_cfg) { c.threads = C.int32_t(n) } }
func GatherStats() Arg   { return func(d *Driver, _ *C.gk_cfg) { d.gatherStats = true } }

func New(args ...Arg) (*Driver, error) {
	d := &Driver{}
	var cfg C.gk_cfg
	for _, a := range args {
		a(d, &cfg)
	}
	var cerr *C.char
	d.e = C.gk_engine_create(&cfg, &cerr)
	if d.e == nil {
		return nil, takeErr(cerr) // no CUDA device => error: there is no CPU fallback
	}
	return d, nil
}

func takeErr(c *C.char) error {
	if c == nil {
		return errors.New("gpudriver: unknown error")
	}
	defer C.gk_free_str(c)
	return errors.New(C.GoString(c))
}

func (d *Driver) Name() string { return Name }

// AddTemplate lowers the template's Rego ahead of time; unsupported constructs are an error here, exactly where a
// Rego compile error surfaces in the reference (pkg/controller/constrainttemplate/constrainttemplate_controller.go:476-479).
func (d *Driver) AddTemplate(_ context.Context, ct *templates.ConstraintTemplate) error {
	src, libs, err := regoSource(ct)
	if err != nil {
		return err
	}
	d.mux.Lock()
	defer d.mux.Unlock()
	kind := C.CString(ct.Spec.CRD.Spec.Names.Kind)
	defer C.free(unsafe.Pointer(kind))
	csrc := C.CString(src)
	defer C.free(unsafe.Pointer(csrc))
	var cerr *C.char
	if len(libs) == 0 {
		if rc := C.gk_add_template(d.e, kind, csrc, C.size_t(len(src)), &cerr); rc != 0 {
			return takeErr(cerr)
		}
		return nil
	}
	// the pointer array lives in C memory (cgo forbids Go memory that holds C pointers... and vice versa)
	n := len(libs)
	arr := (*[1 << 20]*C.char)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	lens := (*[1 << 20]C.size_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.size_t(0)))))
	defer C.free(unsafe.Pointer(arr))
	defer C.free(unsafe.Pointer(lens))
	for i, l := range libs {
		arr[i] = C.CString(l)
		lens[i] = C.size_t(len(l))
		defer C.free(unsafe.Pointer(arr[i]))
	}
	if rc := C.gk_add_template_libs(d.e, kind, csrc, C.size_t(len(src)), (**C.char)(unsafe.Pointer(arr)), (*C.size_t)(unsafe.Pointer(lens)), C.size_t(n), &cerr); rc != 0 {
		return takeErr(cerr)
	}
	return nil
}

func (d *Driver) RemoveTemplate(_ context.Context, ct *templates.ConstraintTemplate) error {
	d.mux.Lock()
	defer d.mux.Unlock()
	kind := C.CString(ct.Spec.CRD.Spec.Names.Kind)
	defer C.free(unsafe.Pointer(kind))
	C.gk_remove_template(d.e, kind)
	return nil
}

// AddConstraint hands the whole constraint (spec.match, spec.parameters, enforcement actions) to the engine: the match
// block becomes the in-kernel pre-filter (pkg/target/target.go:239-254, pkg/mutation/match/match.go:32-65).
func (d *Driver) AddConstraint(_ context.Context, c *unstructured.Unstructured) error {
	raw, err := c.MarshalJSON()
	if err != nil {
		return err
	}
	d.mux.Lock()
	defer d.mux.Unlock()
	var cerr *C.char
	if rc := C.gk_add_constraint(d.e, (*C.char)(unsafe.Pointer(&raw[0])), C.size_t(len(raw)), &cerr); rc != 0 {
		return takeErr(cerr)
	}
	return nil
}

func (d *Driver) RemoveConstraint(_ context.Context, c *unstructured.Unstructured) error {
	d.mux.Lock()
	defer d.mux.Unlock()
	kind, name := C.CString(c.GetKind()), C.CString(c.GetName())
	defer C.free(unsafe.Pointer(kind))
	defer C.free(unsafe.Pointer(name))
	C.gk_remove_constraint(d.e, kind, name)
	return nil
}

// AddData: only Namespace objects matter (namespaceSelector table); paths per pkg/target/target.go:60-66.
func (d *Driver) AddData(_ context.Context, _ string, path storage.Path, data interface{}) error {
	if len(path) >= 4 && path[0] == "cluster" && path[2] == "Namespace" {
		raw, err := json.Marshal(data)
		if err != nil {
			return err
		}
		d.mux.Lock()
		defer d.mux.Unlock()
		name := C.CString(path[3])
		defer C.free(unsafe.Pointer(name))
		var cerr *C.char
		if rc := C.gk_put_namespace(d.e, name, (*C.char)(unsafe.Pointer(&raw[0])), C.size_t(len(raw)), &cerr); rc != 0 {
			return takeErr(cerr)
		}
	}
	return nil
}

func (d *Driver) RemoveData(_ context.Context, _ string, path storage.Path) error {
	if len(path) >= 4 && path[0] == "cluster" && path[2] == "Namespace" {
		d.mux.Lock()
		defer d.mux.Unlock()
		name := C.CString(path[3])
		defer C.free(unsafe.Pointer(name))
		C.gk_remove_namespace(d.e, name)
	}
	return nil
}

// ARGetter / IsAdmissionGetter: how drivers reach the unexported *gkReview -- pkg/drivers/k8scel/driver.go:265-271.
type ARGetter interface {
	GetAdmissionRequest() *admissionv1.AdmissionRequest
}

// Query: one review, the constraints Client.Review already matched (pkg/drivers/k8scel/driver.go:161-250).  The
// webhook path funnels concurrent Query calls through a coalescer (not shown) so that 64 reviews share one launch.
func (d *Driver) Query(ctx context.Context, target string, constraints []*unstructured.Unstructured, review interface{},
	opts ...reviews.ReviewOpt) (*drivers.QueryResponse, error) {
	cfg := &reviews.ReviewCfg{}
	for _, o := range opts {
		o(cfg)
	}
	arGetter, ok := review.(ARGetter)
	if !ok {
		return nil, errors.New("cannot convert review to ARGetter")
	}
	ar := arGetter.GetAdmissionRequest()
	results, stats, err := d.reviewBatch(ctx, []*admissionv1.AdmissionRequest{ar}, []map[string]interface{}{cfg.Namespace},
		cfg.EnforcementPoint, true)
	if err != nil {
		return nil, err
	}
	want := map[string]*unstructured.Unstructured{}
	for _, c := range constraints {
		want[c.GetKind()+"/"+c.GetName()] = c
	}
	out := &drivers.QueryResponse{}
	for _, r := range results {
		if c, ok := want[r.key]; ok {
			out.Results = append(out.Results, &types.Result{Target: target, Msg: r.msg,
				Metadata: map[string]interface{}{"details": r.details}, Constraint: c})
		}
	}
	if d.gatherStats || cfg.StatsEnabled {
		out.StatsEntries = stats
	}
	return out, nil
}

// BatchReviewer is the additive entry point pkg/audit type-asserts for at manager.go:622/:720 (SURVEY.md 8(b)).
type BatchReviewer interface {
	ReviewBatch(ctx context.Context, target string, objs []*unstructured.Unstructured, namespaces []map[string]interface{},
		enforcementPoint string) ([]BatchResult, error)
}

type BatchResult struct {
	Object            int
	Constraint        string // "Kind/name"
	Msg               string
	Details           interface{}
	EnforcementAction string
	ScopedActions     []string
	Autoreject        bool
}

type rawResult struct {
	key, msg string
	details  interface{}
}

func (d *Driver) reviewBatch(_ context.Context, ars []*admissionv1.AdmissionRequest, nss []map[string]interface{}, ep string,
	materialize bool) ([]rawResult, []*instrumentation.StatsEntry, error) {
	d.mux.RLock()
	defer d.mux.RUnlock()
	n := len(ars)
	objs := make([]C.gk_obj, n)
	keep := make([][]byte, 0, 3*n) // keeps Go memory alive and pinned for the duration of the call
	for i, ar := range ars {
		o := &objs[i]
		if ar.Object.Raw != nil {
			o.json, o.len = (*C.char)(unsafe.Pointer(&ar.Object.Raw[0])), C.size_t(len(ar.Object.Raw))
		}
		if ar.OldObject.Raw != nil {
			o.old_json, o.old_len = (*C.char)(unsafe.Pointer(&ar.OldObject.Raw[0])), C.size_t(len(ar.OldObject.Raw))
		}
		if nss[i] != nil {
			b, _ := json.Marshal(nss[i])
			keep = append(keep, b)
			o.ns_json, o.ns_len = (*C.char)(unsafe.Pointer(&b[0])), C.size_t(len(b))
		}
		o.source = C.GK_SOURCE_ORIGINAL
	}
	cep := C.CString(ep)
	defer C.free(unsafe.Pointer(cep))
	var res C.gk_result
	var cerr *C.char
	flags := C.uint32_t(0)
	if materialize {
		flags = C.GK_F_MATERIALIZE
	}
	if rc := C.gk_review_batch(d.e, &objs[0], C.size_t(n), cep, flags, &res, &cerr); rc != 0 {
		return nil, nil, takeErr(cerr)
	}
	defer C.gk_free_result(&res)
	_ = keep
	out := make([]rawResult, 0, int(res.n_violations))
	vs := unsafe.Slice(res.violations, int(res.n_violations))
	for _, v := range vs {
		var details interface{}
		if dj := C.GoString(v.details_json); dj != "" {
			_ = json.Unmarshal([]byte(dj), &details)
		}
		out = append(out, rawResult{key: C.GoString(C.gk_result_constraint_key(&res, v.constraint)), msg: C.GoString(v.msg), details: details})
	}
	stats := []*instrumentation.StatsEntry{{Scope: "batch", StatsFor: fmt.Sprintf("%d reviews", n),
		Stats: []*instrumentation.Stat{
			{Name: "kernelTimeNS", Value: uint64(float64(res.kernel_ms) * 1e6), Source: instrumentation.Source{Type: instrumentation.EngineSourceType, Value: Name}},
			{Name: "flattenTimeNS", Value: uint64(float64(res.flatten_ms) * 1e6), Source: instrumentation.Source{Type: instrumentation.EngineSourceType, Value: Name}},
			{Name: "batchSize", Value: n, Source: instrumentation.Source{Type: instrumentation.EngineSourceType, Value: Name}},
			{Name: "bytesRead", Value: uint64(res.alg_bytes), Source: instrumentation.Source{Type: instrumentation.EngineSourceType, Value: Name}},
		}}}
	return out, stats, nil
}

func (d *Driver) Dump(_ context.Context) (string, error) {
	d.mux.RLock()
	defer d.mux.RUnlock()
	c := C.gk_dump(d.e)
	defer C.gk_free_str(c)
	return C.GoString(c), nil
}

func (d *Driver) GetDescriptionForStat(statName string) (string, error) {
	cs := C.CString(statName)
	defer C.free(unsafe.Pointer(cs))
	if desc := C.gk_stat_description(cs); desc != nil {
		return C.GoString(desc), nil
	}
	return "", fmt.Errorf("unknown stat name for Rego (GPU): %s", statName)
}

// regoSource returns the entry-point module and the template's libs (`package lib.<...>` modules the entry point imports as
// data.lib.<...>): `code[engine=Rego].source.{rego,libs}` first, the legacy `rego` / `libs` target fields otherwise.
func regoSource(ct *templates.ConstraintTemplate) (string, []string, error) {
	if len(ct.Spec.Targets) != 1 {
		return "", nil, errors.New("expected exactly one target")
	}
	t := ct.Spec.Targets[0]
	for _, code := range t.Code {
		if code.Engine == Name {
			if m, ok := code.Source.Value.(map[string]interface{}); ok {
				if s, ok := m["rego"].(string); ok {
					var libs []string
					if ls, ok := m["libs"].([]interface{}); ok {
						for _, l := range ls {
							if str, ok := l.(string); ok {
								libs = append(libs, str)
							}
						}
					}
					return s, libs, nil
				}
			}
		}
	}
	if t.Rego != "" {
		return t.Rego, t.Libs, nil
	}
	return "", nil, errors.New("no Rego source in template (ErrNoDriver)")
}
