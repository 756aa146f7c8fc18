// NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no Go toolchain): see README.md in this directory.
//
// Drop into a checkout of github.com/everpeace/kube-throttler (e.g. as bench/throttle_bench_test.go) and run
//   go test ./bench -run xxx -bench . -benchtime 3x
// It drives the reference's own v1alpha1 types: ThrottleSelector.MatchesToPod (throttle_selector.go:30-54) and
// Throttle.CheckThrottledFor (throttle_types.go:128-153) for every (pending pod, throttle of its namespace), which is
// the inner loop of ThrottleController.CheckThrottled (throttle_controller.go:349-397).
package bench

import (
	"fmt"
	"math/rand"
	"runtime"
	"sync"
	"testing"

	"github.com/everpeace/kube-throttler/pkg/apis/schedule/v1alpha1"
	corev1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/api/resource"
	metav1 "k8s.io/apimachinery/pkg/apis/meta/v1"
)

const (
	nThrottles  = 1000
	nNamespaces = 50
	nPending    = 10000
	nKeys       = 16
	nValues     = 64
)

func mkPod(rng *rand.Rand, i int) *corev1.Pod {
	labels := map[string]string{}
	for len(labels) < 2+rng.Intn(7) {
		labels[fmt.Sprintf("k%d", rng.Intn(nKeys))] = fmt.Sprintf("v%d", rng.Intn(nValues))
	}
	return &corev1.Pod{
		ObjectMeta: metav1.ObjectMeta{Namespace: fmt.Sprintf("ns%d", rng.Intn(nNamespaces)), Name: fmt.Sprintf("p%d", i), Labels: labels},
		Spec: corev1.PodSpec{Containers: []corev1.Container{{Name: "c", Resources: corev1.ResourceRequirements{Requests: corev1.ResourceList{
			corev1.ResourceCPU:    *resource.NewMilliQuantity(int64(50*(1+rng.Intn(80))), resource.DecimalSI),
			corev1.ResourceMemory: *resource.NewQuantity(int64(64<<20)<<uint(rng.Intn(9)), resource.BinarySI),
		}}}}},
	}
}

func mkThrottle(rng *rand.Rand, i int) v1alpha1.Throttle {
	thr := v1alpha1.Throttle{ObjectMeta: metav1.ObjectMeta{Namespace: fmt.Sprintf("ns%d", i%nNamespaces), Name: fmt.Sprintf("t%d", i)}}
	thr.Spec.ThrottlerName = "kube-throttler"
	thr.Spec.Threshold = v1alpha1.ResourceAmount{
		ResourceCounts:   &v1alpha1.ResourceCounts{Pod: 50 + rng.Intn(100)},
		ResourceRequests: corev1.ResourceList{corev1.ResourceCPU: resource.MustParse(fmt.Sprintf("%d", 20+rng.Intn(200)))},
	}
	thr.Spec.Selector = v1alpha1.ThrottleSelector{SelecterTerms: []v1alpha1.ThrottleSelectorTerm{{
		PodSelector: metav1.LabelSelector{MatchLabels: map[string]string{fmt.Sprintf("k%d", rng.Intn(nKeys)): fmt.Sprintf("v%d", rng.Intn(nValues))}},
	}}}
	thr.Status.Used = v1alpha1.ResourceAmount{
		ResourceCounts:   &v1alpha1.ResourceCounts{Pod: rng.Intn(120)},
		ResourceRequests: corev1.ResourceList{corev1.ResourceCPU: resource.MustParse(fmt.Sprintf("%d", rng.Intn(200)))},
	}
	return thr
}

func BenchmarkCheckThrottled(b *testing.B) {
	rng := rand.New(rand.NewSource(2))
	byNs := map[string][]v1alpha1.Throttle{}
	for i := 0; i < nThrottles; i++ {
		t := mkThrottle(rng, i)
		byNs[t.Namespace] = append(byNs[t.Namespace], t)
	}
	pods := make([]*corev1.Pod, nPending)
	for i := range pods {
		pods[i] = mkPod(rng, i)
	}
	workers := runtime.GOMAXPROCS(0)
	b.ResetTimer()
	for it := 0; it < b.N; it++ {
		var wg sync.WaitGroup
		for w := 0; w < workers; w++ {
			wg.Add(1)
			go func(w int) {
				defer wg.Done()
				for i := w; i < len(pods); i += workers {
					pod := pods[i]
					for _, thr := range byNs[pod.Namespace] { // affectedThrottles: lister by namespace, then the selector
						match, err := thr.Spec.Selector.MatchesToPod(pod)
						if err != nil || !match {
							continue
						}
						_ = thr.CheckThrottledFor(pod, v1alpha1.ResourceAmount{}, false)
					}
				}
			}(w)
		}
		wg.Wait()
	}
	pairs := float64(nPending) * float64(nThrottles/nNamespaces) * float64(b.N)
	b.ReportMetric(pairs/b.Elapsed().Seconds(), "pod_x_throttle_in_namespace_checks/s")
	b.ReportMetric(float64(nPending)*float64(nThrottles)*float64(b.N)/b.Elapsed().Seconds(), "all_pairs_equivalent_checks/s")
}
