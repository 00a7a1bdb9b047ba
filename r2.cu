__global__ void k(float* p, float a, float b){
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};"::"l"(p),"f"(a),"f"(b):"memory");
}
