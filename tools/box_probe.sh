#!/bin/bash
# What the GPU box is: host cores / NUMA / PCIe topology / memlock / huge pages, and whether any
# GStreamer (libgstvideo-1.0: the library that holds gst_video_blend) can be found on it.
#   gpurun -- 'bash tools/box_probe.sh > gpurun_out/box_probe.txt 2>&1'
echo "== date"; date -u
echo "== nproc / lscpu"; nproc; lscpu 2>/dev/null | grep -Ei "model name|socket|core|thread|numa|^cpu\(s\)|l3|hypervisor|virtual"
echo "== numa nodes"; ls -d /sys/devices/system/node/node* 2>/dev/null
for n in /sys/devices/system/node/node*; do echo "$n cpulist: $(cat $n/cpulist 2>/dev/null)"; grep -E "MemTotal|MemFree" $n/meminfo 2>/dev/null; done
echo "== meminfo"; grep -Ei "memtotal|memfree|hugepages|hugepagesize" /proc/meminfo
echo "== THP"; cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null
echo "== ulimit -l"; ulimit -l
echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
echo "== affinity of this shell"; taskset -p $$ 2>/dev/null || grep Cpus_allowed_list /proc/self/status
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv
echo "== topo"; nvidia-smi topo -m 2>&1
echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d class=$(cat $d/class) numa=$(cat $d/numa_node 2>/dev/null) local_cpulist=$(cat $d/local_cpulist 2>/dev/null)"; fi; done
echo "== iommu"; ls /sys/kernel/iommu_groups 2>/dev/null | wc -l; cat /proc/cmdline 2>/dev/null
echo "== gstreamer probe"
for t in gst-launch-1.0 gst-inspect-1.0 pkg-config meson; do printf "%s: " $t; command -v $t || echo absent; done
ldconfig -p 2>/dev/null | grep -iE "gst|glib-2|cairo|pixman|pango" || echo "ldconfig: no gst/glib/cairo/pixman/pango library"
echo "-- find libgst* / libglib* / libcairo* / libpixman*"
find / -xdev \( -name 'libgst*' -o -name 'libglib-2*' -o -name 'libcairo*' -o -name 'libpixman*' -o -name 'video-blend*' -o -name 'gstvideo*' \) 2>/dev/null | head -50
echo "-- python gi / cairo"
python -c "import gi; print('gi', gi.__version__)" 2>&1 | tail -1
python -c "import cairo; print('cairo', cairo.version)" 2>&1 | tail -1
python - <<'P'
import importlib.util
for m in ("cv2","av","imageio_ffmpeg","PIL","skia","cairocffi","pgi"):
    print(m, "present" if importlib.util.find_spec(m) else "absent")
try:
    import cv2
    bi = cv2.getBuildInformation()
    for l in bi.splitlines():
        if "GStreamer" in l or "FFMPEG" in l: print(l.strip())
except Exception as e: print("cv2:", e)
P
echo "-- network"; (timeout 5 python - <<'P'
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=3); print("network: reachable")
except Exception as e: print("network: none (", e, ")")
P
) 2>&1
echo "-- apt"; (timeout 10 apt-get download libgstreamer-plugins-base1.0-0 2>&1 | tail -2) || echo "apt-get download: failed/timeout"
ls /var/cache/apt/archives/*.deb 2>/dev/null | head
